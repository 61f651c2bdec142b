// session.cu -- a-10: the frame loop with HOST buffers (fast_artistic_video_core.lua:189-229 +
// fast_artistic_video.lua:93-170).  One session = one H x W stream on one GPU.
//
// The recurrent state last_frame_stylized stays on the device as unclamped fp32 (fast_artistic_video.lua:169).
// Three streams: H2D (frame i+1 inputs), compute (frame i), D2H (frame i-1 result); inputs and outputs are
// double buffered and ordered with events, so copies overlap compute whenever the caller's host buffers are
// pinned.  Nothing in the loop synchronises the host except fav_session_sync(); host buffers handed to a call must stay
// untouched until then (the copies are asynchronous).  Per frame the compute stream runs TWO launches: the fused
// temporal-stage kernel (fav_run_next_image_flows) and the CUDA graph of the network.
// The flows variant evaluates checkConsistency in its 3-argument mode (video_dataset/make_occlusions.sh:31-36); the
// 4-argument structure term of makeOptFlow_deepflow.sh:59-60 is available through fav_compute_corners +
// fav_consistency_check + the cert-given variant.
#include <memory>
#include <vector>

#include "fav_common.cuh"


using namespace fav;

struct fav_session {
  fav_net_t *net = nullptr;
  fav_net_t *net_img = nullptr;  // optional separate image model for single images (-model_img, core.lua:61-68,146)
  int H = 0, W = 0;
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  struct InSet {
    float *content = nullptr, *flow = nullptr, *flow_fw = nullptr, *cert_raw = nullptr, *cert = nullptr;
    cudaEvent_t uploaded = nullptr, consumed = nullptr;
    bool used = false;
  } in[2];
  float *out[2] = {nullptr, nullptr};
  cudaEvent_t computed[2] = {nullptr, nullptr}, downloaded[2] = {nullptr, nullptr};
  static constexpr int kDoneRing = 64;  // per-frame completion events for host threads (fav_session_frame_done)
  cudaEvent_t done[kDoneRing] = {};
  bool out_used[2] = {false, false};
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  uint64_t frame = 0;
  bool have_prev = false;
  std::vector<void *> allocs;
};

static int s_alloc(fav_session *s, float **p, size_t n) {
  void *d = nullptr;
  FAV_TRY(check_cuda(cudaMalloc(&d, n * sizeof(float)), "cudaMalloc(session)"));
  s->allocs.push_back(d);
  *p = (float *)d;
  return FAV_OK;
}

extern "C" {

void fav_session_destroy(fav_session_t *s) {
  if (!s) return;
  cudaDeviceSynchronize();
  for (void *p : s->allocs) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (s->in[i].uploaded) cudaEventDestroy(s->in[i].uploaded);
    if (s->in[i].consumed) cudaEventDestroy(s->in[i].consumed);
    if (s->computed[i]) cudaEventDestroy(s->computed[i]);
    if (s->downloaded[i]) cudaEventDestroy(s->downloaded[i]);
  }
  for (cudaEvent_t e : s->done)
    if (e) cudaEventDestroy(e);
  if (s->t0) cudaEventDestroy(s->t0);
  if (s->t1) cudaEventDestroy(s->t1);
  if (s->s_h2d) cudaStreamDestroy(s->s_h2d);
  if (s->s_comp) cudaStreamDestroy(s->s_comp);
  if (s->s_d2h) cudaStreamDestroy(s->s_d2h);
  delete s;
}

int fav_session_create(fav_net_t *net, int H, int W, fav_session_t **out) {
  FAV_REQUIRE(net && out, "fav_session_create: null argument");
  FAV_REQUIRE(H >= 16 && W >= 16 && H % 4 == 0 && W % 4 == 0,
              "fav_session_create: frame size %dx%d: H and W must be multiples of 4 and >= 16 (reflect-start nets restore the "
              "input size only then)", W, H);
  FAV_TRY(require_device());
  std::unique_ptr<fav_session, void (*)(fav_session *)> s(new fav_session(), fav_session_destroy);
  s->net = net; s->H = H; s->W = W;
  FAV_TRY(check_cuda(cudaStreamCreateWithFlags(&s->s_h2d, cudaStreamNonBlocking), "cudaStreamCreate"));
  FAV_TRY(check_cuda(cudaStreamCreateWithFlags(&s->s_comp, cudaStreamNonBlocking), "cudaStreamCreate"));
  FAV_TRY(check_cuda(cudaStreamCreateWithFlags(&s->s_d2h, cudaStreamNonBlocking), "cudaStreamCreate"));
  const size_t HW = (size_t)H * W;
  for (int i = 0; i < 2; ++i) {
    FAV_TRY(s_alloc(s.get(), &s->in[i].content, 3 * HW));
    FAV_TRY(s_alloc(s.get(), &s->in[i].flow, 2 * HW));
    FAV_TRY(s_alloc(s.get(), &s->in[i].flow_fw, 2 * HW));
    FAV_TRY(s_alloc(s.get(), &s->in[i].cert_raw, HW));
    FAV_TRY(s_alloc(s.get(), &s->in[i].cert, HW));
    FAV_TRY(s_alloc(s.get(), &s->out[i], 3 * HW));
    FAV_TRY(check_cuda(cudaEventCreateWithFlags(&s->in[i].uploaded, cudaEventDisableTiming), "cudaEventCreate"));
    FAV_TRY(check_cuda(cudaEventCreateWithFlags(&s->in[i].consumed, cudaEventDisableTiming), "cudaEventCreate"));
    FAV_TRY(check_cuda(cudaEventCreateWithFlags(&s->computed[i], cudaEventDisableTiming), "cudaEventCreate"));
    FAV_TRY(check_cuda(cudaEventCreateWithFlags(&s->downloaded[i], cudaEventDisableTiming), "cudaEventCreate"));
  }
  for (int i = 0; i < fav_session::kDoneRing; ++i)  // waited on by host threads: block, do not spin
    FAV_TRY(check_cuda(cudaEventCreateWithFlags(&s->done[i], cudaEventDisableTiming | cudaEventBlockingSync), "cudaEventCreate"));
  FAV_TRY(check_cuda(cudaEventCreate(&s->t0), "cudaEventCreate"));
  FAV_TRY(check_cuda(cudaEventCreate(&s->t1), "cudaEventCreate"));
  *out = s.release();
  return FAV_OK;
}

// mode 0: first frame; 1: cert given; 2: cert from the flow pair
static int session_step(fav_session *s, int mode, const float *content_host, const float *flow_a, const float *flow_b,
                        const float *cert_host, int min_filter_r, int border_mode, float *out_host) {
  FAV_REQUIRE(s && content_host && out_host, "fav_session: null argument");
  FAV_REQUIRE(mode == 0 || s->have_prev, "fav_session_run_next_image: no previous frame (call run_image first)");
  FAV_REQUIRE(min_filter_r == 0 || ((min_filter_r & 1) && min_filter_r <= 15), "occlusions_min_filter must be odd <= 15");
  const int H = s->H, W = s->W;
  const size_t HW = (size_t)H * W;
  const int si = (int)(s->frame & 1), so = (int)(s->frame & 1);
  fav_session::InSet &in = s->in[si];
  // ---- H2D (waits until the frame that last used this input set has consumed it)
  if (in.used) FAV_TRY(check_cuda(cudaStreamWaitEvent(s->s_h2d, in.consumed, 0), "cudaStreamWaitEvent"));
  FAV_TRY(check_cuda(cudaMemcpyAsync(in.content, content_host, 3 * HW * 4, cudaMemcpyHostToDevice, s->s_h2d), "H2D content"));
  if (mode == 1) {
    FAV_TRY(check_cuda(cudaMemcpyAsync(in.flow, flow_a, 2 * HW * 4, cudaMemcpyHostToDevice, s->s_h2d), "H2D flow"));
    FAV_TRY(check_cuda(cudaMemcpyAsync(in.cert_raw, cert_host, HW * 4, cudaMemcpyHostToDevice, s->s_h2d), "H2D cert"));
  } else if (mode == 2) {
    // backward flow arrives in .flo order (u,v); the warp wants (dy,dx) = (v,u): swap planes while uploading
    FAV_TRY(check_cuda(cudaMemcpyAsync(in.flow + HW, flow_a, HW * 4, cudaMemcpyHostToDevice, s->s_h2d), "H2D flow u"));
    FAV_TRY(check_cuda(cudaMemcpyAsync(in.flow, flow_a + HW, HW * 4, cudaMemcpyHostToDevice, s->s_h2d), "H2D flow v"));
    FAV_TRY(check_cuda(cudaMemcpyAsync(in.flow_fw, flow_b, 2 * HW * 4, cudaMemcpyHostToDevice, s->s_h2d), "H2D flow fw"));
  }
  FAV_TRY(check_cuda(cudaEventRecord(in.uploaded, s->s_h2d), "cudaEventRecord"));
  // ---- compute
  FAV_TRY(check_cuda(cudaStreamWaitEvent(s->s_comp, in.uploaded, 0), "cudaStreamWaitEvent"));
  if (s->out_used[so])  // the D2H of the frame that last wrote out[so] must be finished
    FAV_TRY(check_cuda(cudaStreamWaitEvent(s->s_comp, s->downloaded[so], 0), "cudaStreamWaitEvent"));
  FAV_TRY(check_cuda(cudaEventRecord(s->t0, s->s_comp), "cudaEventRecord"));
  if (mode == 0) {
    FAV_TRY(fav_run_image(s->net_img ? s->net_img : s->net, in.content, nullptr, H, W, s->out[so], s->s_comp));
  } else {
    // ONE temporal-stage kernel per frame: occlusion test (mode 2) or given certainty (mode 1) -> min filter -> warp ->
    // preprocess -> mask -> concat -> first operand of the net; then one graph launch
    FAV_TRY(fav_run_next_image_flows(s->net, in.content, s->out[so ^ 1], in.flow, mode == 2 ? in.flow_fw : nullptr,
                                     mode == 1 ? in.cert_raw : nullptr, nullptr, nullptr, H, W, min_filter_r, border_mode,
                                     s->out[so], s->s_comp));
  }
  FAV_TRY(check_cuda(cudaEventRecord(s->t1, s->s_comp), "cudaEventRecord"));
  FAV_TRY(check_cuda(cudaEventRecord(in.consumed, s->s_comp), "cudaEventRecord"));
  FAV_TRY(check_cuda(cudaEventRecord(s->computed[so], s->s_comp), "cudaEventRecord"));
  in.used = true;
  // ---- D2H
  FAV_TRY(check_cuda(cudaStreamWaitEvent(s->s_d2h, s->computed[so], 0), "cudaStreamWaitEvent"));
  FAV_TRY(check_cuda(cudaMemcpyAsync(out_host, s->out[so], 3 * HW * 4, cudaMemcpyDeviceToHost, s->s_d2h), "D2H out"));
  FAV_TRY(check_cuda(cudaEventRecord(s->downloaded[so], s->s_d2h), "cudaEventRecord"));
  FAV_TRY(check_cuda(cudaEventRecord(s->done[s->frame % fav_session::kDoneRing], s->s_d2h), "cudaEventRecord"));
  s->out_used[so] = true;
  s->have_prev = true;
  s->frame++;
  return FAV_OK;
}

int fav_session_set_image_model(fav_session_t *s, fav_net_t *net_img) {
  FAV_REQUIRE(s, "fav_session_set_image_model: null session");
  s->net_img = net_img;
  return FAV_OK;
}

int fav_session_run_image(fav_session_t *s, const float *content_host, float *out_host) {
  return session_step(s, 0, content_host, nullptr, nullptr, nullptr, 0, 0, out_host);
}

int fav_session_run_next_image(fav_session_t *s, const float *content_host, const float *flow_host,
                               const float *cert_host, int min_filter_r, int border_mode, float *out_host) {
  FAV_REQUIRE(flow_host && cert_host, "fav_session_run_next_image: null flow / cert");
  return session_step(s, 1, content_host, flow_host, nullptr, cert_host, min_filter_r, border_mode, out_host);
}

int fav_session_run_next_image_flows(fav_session_t *s, const float *content_host, const float *flow_bw_uv_host,
                                     const float *flow_fw_uv_host, int min_filter_r, int border_mode, float *out_host) {
  FAV_REQUIRE(flow_bw_uv_host && flow_fw_uv_host, "fav_session_run_next_image_flows: null flow");
  return session_step(s, 2, content_host, flow_bw_uv_host, flow_fw_uv_host, nullptr, min_filter_r, border_mode, out_host);
}

// the stylized frame of call number `frame_index` (0-based count of run_* calls on this session) has landed in its out_host
// buffer; wait != 0 blocks until then.  One completion event per frame in a ring of 64: frames older than that have landed
// long ago (the ring slot then belongs to a later frame, and waiting for it is never too short).
int fav_session_frame_done(fav_session_t *s, uint64_t frame_index, int wait) {
  FAV_REQUIRE(s && frame_index < s->frame, "fav_session_frame_done: frame %llu was never enqueued", (unsigned long long)frame_index);
  cudaEvent_t ev = s->done[frame_index % fav_session::kDoneRing];
  if (wait) return check_cuda(cudaEventSynchronize(ev), "cudaEventSynchronize(frame done)");
  cudaError_t e = cudaEventQuery(ev);
  if (e == cudaSuccess) return FAV_OK;
  if (e == cudaErrorNotReady) { set_error("frame %llu still in flight", (unsigned long long)frame_index); return FAV_ERR_INVALID; }
  return check_cuda(e, "cudaEventQuery(frame done)");
}

int fav_session_sync(fav_session_t *s) {
  FAV_REQUIRE(s, "fav_session_sync: null session");
  FAV_TRY(check_cuda(cudaStreamSynchronize(s->s_h2d), "sync h2d"));
  FAV_TRY(check_cuda(cudaStreamSynchronize(s->s_comp), "sync compute"));
  return check_cuda(cudaStreamSynchronize(s->s_d2h), "sync d2h");
}

float fav_session_last_gpu_ms(fav_session_t *s) {
  if (!s || !s->t0 || !s->t1) return -1.f;
  if (cudaEventSynchronize(s->t1) != cudaSuccess) return -1.f;
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, s->t0, s->t1) != cudaSuccess) return -1.f;
  return ms;
}
}
