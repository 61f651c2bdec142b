// net.cu -- host side of the stylization network: arch parser, parameter store, weight repacking, per-size
// execution plan, forward, run_image / run_next_image.
//   models_video.build_model              fast_artistic_video/models_video.lua:55-140
//   build_res_block (reflect-start)       models_video.lua:41-53
//   lazily inserted reflection padding    train_video.lua:316-325
//   run_image / run_next_image            fast_artistic_video_core.lua:121-180
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>

#include "conv_plan.hpp"
#include "conv_res_plan.hpp"

namespace fav {

int launch_temporal_input(const float *content, const float *prev, const float *flow, const float *cert,
                          const float *fill, const float *flow_mask, float *out7, int H, int W, int border_mode,
                          bool first, cudaStream_t st);
int launch_temporal_stage(const float *content, const float *prev, const float *flow, const float *fw_uv, const float *cert_raw,
                          const float *fill, const float *flow_mask, float *out7, float *cert_out, const Operand *dst, int R,
                          int H, int W, int r, int border_mode, cudaStream_t st);
int launch_min_filter(const float *in, float *out, int n, int H, int W, int r, cudaStream_t st);
int launch_consistency(const float *f1u, const float *f1v, const float *f2u, const float *f2v, const float *structure,
                       const float *avg_dev, float avg_host, uint8_t *rel, float *cert, int W, int H, cudaStream_t st);
int launch_temporal_input_packed(const float *content, const float *prev, const float *flow, const float *cert,
                                 const float *fill, const float *flow_mask, const Operand &dst, int R, int H, int W,
                                 int border_mode, bool first, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
struct Param {
  std::string name;
  int64_t shape[4] = {1, 1, 1, 1};
  int64_t numel = 0;
  std::vector<float> host;
  bool set = false;
};

struct InDef {
  std::string name;
  int C = 0;
  float *d_gamma = nullptr, *d_beta = nullptr;
  int pw = -1, pb = -1;
};

// arch-level program
struct SpecOp {
  int kind;  // 0 conv(+in+relu), 1 conv block: two 3x3 convs + IN (RX: + ShaveImage / Identity skip; CX: no skip, ReLU after),
             // 2 nearest upsampling (+in+relu)
  bool skip = true;   // kind 1: residual (RX) or plain (CX) block
  int shave = 2;      // kind 1: ShaveImage(2) ('reflect-start') or Identity ('zero') on the skip branch
  int scale = 1;
  int conv[2] = {-1, -1};
  int inorm[2] = {-1, -1};
  bool relu = false, last = false;
};

struct PlanStep {
  int kind;  // 0 conv, 1 instance norm (+relu, +skip) -> operand, 2 nearest upsampling + instance norm + relu
  int scale = 1;
  int conv = -1, inorm = -1;
  int src = -1, dst = -1, skip = -1;
  int relu = 0;
  int shave = 2;
  RawTensor raw;
  std::vector<ConvJob> tc;
  std::vector<ResJob> res;   // non-empty: the tcgen05 path of this step is conv_res.cu (raw output planar)
  std::vector<SimtJob> simt;
  int stats_off = 0;
  int layer_index = -1;  // arch token index whose output this step completes (for fav_net_layer_output)
  bool fused_nl = false; // kind 1: the tcgen05 consumer normalises on load (conv_tc.cu), this pass is skipped
};

struct Plan {
  int H = 0, W = 0;
  std::vector<Operand> ops;
  std::vector<void *> allocs;
  float *raw_buf = nullptr, *raw_buf2 = nullptr;  // consecutive convolutions alternate (norm-on-load reads one, writes the other)
  double *stats = nullptr;
  size_t stats_bytes = 0;
  float *msb = nullptr;
  float *in7 = nullptr, *out3 = nullptr;  // scratch for run_image / run_next_image
  float *cert_a = nullptr, *cert_b = nullptr;  // scratch of the unfused fallback of fav_run_next_image_flows
  std::vector<PlanStep> steps;
  std::map<int, int> layer_operand;  // arch token index -> operand holding its output
  // CUDA graphs of the whole per-frame kernel sequence (run_[next_]image), one per destination buffer: a frame is ~40
  // launches and the host needs ~1.1 ms to enqueue them one by one (measured, bench e2e), about the GPU time itself
  struct GraphEntry { float *out3; cudaGraphExec_t exec; uint64_t launches, last_use; };
  std::vector<GraphEntry> graphs;
  uint64_t use_clock = 0;
  int eager_runs = 0;
  ~Plan() {
    for (GraphEntry &g : graphs) cudaGraphExecDestroy(g.exec);
    for (void *p : allocs) cudaFree(p);
  }
};

}  // namespace fav

using namespace fav;

struct fav_net {
  std::string arch;
  std::string padding_type = "reflect-start";
  float tanh_c = 150.f;
  int in_dim = 7;
  int reflect_pad = 0;
  std::vector<Param> params;
  std::vector<ConvDef> convs;
  std::vector<InDef> inorms;
  std::vector<SpecOp> ops;
  bool finalized = false;
  int conv_impl = 0;
  int num_sms = 148;
  cudaStream_t cap_stream = nullptr;  // graph capture happens here (the caller's stream may be the legacy default stream)
  int device = 0;
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> plans;
  Plan *last_plan = nullptr;
  std::vector<void *> allocs;
  ~fav_net() {
    plans.clear();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    for (void *p : allocs) cudaFree(p);
  }
};

namespace fav {

static int add_param(fav_net *net, const std::string &name, std::initializer_list<int64_t> shape) {
  Param p;
  p.name = name;
  int i = 0;
  p.numel = 1;
  for (int64_t s : shape) {
    p.shape[i++] = s;
    p.numel *= s;
  }
  net->params.push_back(std::move(p));
  return (int)net->params.size() - 1;
}

static int add_conv(fav_net *net, const std::string &name, int cin, int cout, int k, int stride, int pad, bool tr,
                    int adj) {
  ConvDef c;
  init_conv_def(c, name, cin, cout, k, stride, pad, tr, adj);
  build_phases(c);
  if (tr) c.pw = add_param(net, name + ".weight", {cin, cout, k, k});
  else c.pw = add_param(net, name + ".weight", {cout, cin, k, k});
  c.pb = add_param(net, name + ".bias", {cout});
  net->convs.push_back(std::move(c));
  return (int)net->convs.size() - 1;
}
static int add_in(fav_net *net, const std::string &name, int C) {
  InDef n;
  n.name = name; n.C = C;
  n.pw = add_param(net, name + ".weight", {C});
  n.pb = add_param(net, name + ".bias", {C});
  net->inorms.push_back(std::move(n));
  return (int)net->inorms.size() - 1;
}

template <class T>
static int dev_upload(fav_net *net, const std::vector<T> &h, T **out) {
  void *d = nullptr;
  FAV_TRY(check_cuda(cudaMalloc(&d, std::max<size_t>(h.size() * sizeof(T), 16)), "cudaMalloc(weights)"));
  net->allocs.push_back(d);
  FAV_TRY(check_cuda(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy(weights)"));
  *out = (T *)d;
  return FAV_OK;
}

static int pack_conv_weights(fav_net *net, ConvDef &c) {
  const std::vector<float> &w = net->params[c.pw].host;
  std::vector<float> bias(c.Npad, 0.f);
  for (int i = 0; i < c.cout; ++i) bias[i] = net->params[c.pb].host[i];
  FAV_TRY(dev_upload(net, bias, &c.d_bias));
  for (auto &ph : c.phases) {
    FAV_TRY(build_phase_tables(c, ph));
    // CUDA-core comparator layout: [tap][Cin_pad][Cout_pad8]
    std::vector<float> ws((size_t)ph.taps.size() * c.cin_pad * c.cout_pad8, 0.f);
    for (size_t t = 0; t < ph.taps.size(); ++t)
      for (int ci = 0; ci < c.cin; ++ci)
        for (int co = 0; co < c.cout; ++co)
          ws[(t * c.cin_pad + ci) * c.cout_pad8 + co] = weight_at(c, w, co, ci, ph.taps[t].ky, ph.taps[t].kx);
    FAV_TRY(dev_upload(net, ws, &ph.d_w_simt));
    std::vector<uint16_t> pk = pack_phase_weights(c, ph, w);
    uint16_t *d = nullptr;
    FAV_TRY(dev_upload(net, pk, &d));
    ph.d_b_tc = reinterpret_cast<uint4 *>(d);
  }
  if (c.has_fold) {  // transposed conv: one tcgen05 job for all four sub-pixel phases
    std::vector<uint16_t> pk = pack_phase_fold(c, c.fold, w);
    uint16_t *d = nullptr;
    FAV_TRY(dev_upload(net, pk, &d));
    c.fold.d_b_tc = reinterpret_cast<uint4 *>(d);
  }
  return FAV_OK;
}

// ---- plan ---------------------------------------------------------------------------------------------------
static int alloc_zero(Plan &pl, void **out, size_t bytes) {
  FAV_TRY(check_cuda(cudaMalloc(out, bytes), "cudaMalloc(activations)"));
  pl.allocs.push_back(*out);
  return check_cuda(cudaMemset(*out, 0, bytes), "cudaMemset(activations)");
}

// operand feeding convolution `c` (its pad / parity requirements)
static int make_operand(Plan &pl, int C, int H, int W, const ConvDef *consumer, int *index) {
  Operand o = operand_geometry(C, H, W, consumer);
  FAV_TRY(alloc_zero(pl, (void **)&o.hi, o.elems16 * 16));
  FAV_TRY(alloc_zero(pl, (void **)&o.lo, o.elems16 * 16));
  pl.ops.push_back(o);
  *index = (int)pl.ops.size() - 1;
  return FAV_OK;
}

static int build_conv_jobs(fav_net *net, Plan &pl, PlanStep &st, const ConvDef &c, bool last, float *out3) {
  const Operand &in = pl.ops[st.src];
  bool fold_done = false;
  for (const ConvPhase &ph0 : c.phases) {
    const bool use_fold = c.has_fold && !getenv("FAV_NO_FOLD");
    const ConvPhase &ph = ph0;
    ConvJob j;
    if (use_fold) {
      if (!fold_done) {
        FAV_TRY(fill_conv_job(c, c.fold, in, j));
        j.b = c.fold.d_b_tc; j.bias = c.d_bias;
        j.raw = st.raw.p; j.raw_Cq = st.raw.Cq; j.raw_Wp = st.raw.Wp;
        j.final_mode = 0; j.out3 = out3; j.tanh_c = net->tanh_c;
        conv_tc_choose_slots(j);
        if (const char *e = getenv("FAV_DBG")) j.dbg = atoi(e);
        if (conv_tc_smem_bytes(j) > 227 * 1024) { set_error("conv %s: shared memory budget exceeded (fold)", c.name.c_str()); return FAV_ERR_UNSUPPORTED; }
        st.tc.push_back(j);
        fold_done = true;
      }
    } else {
    FAV_TRY(fill_conv_job(c, ph, in, j));
    j.b = ph.d_b_tc; j.bias = c.d_bias;
    j.raw = st.raw.p; j.raw_Cq = st.raw.Cq; j.raw_Wp = st.raw.Wp;
    j.final_mode = last ? 1 : 0; j.out3 = out3; j.tanh_c = net->tanh_c;
    conv_tc_choose_slots(j);
    if (const char *e = getenv("FAV_DBG")) j.dbg = atoi(e);
    if (conv_tc_smem_bytes(j) > 227 * 1024) {
      set_error("conv %s: shared memory budget exceeded", c.name.c_str());
      return FAV_ERR_UNSUPPORTED;
    }
    st.tc.push_back(j);
    }

    SimtJob s;
    memset(&s, 0, sizeof(s));
    s.in = in; s.sy = s.sx = c.in_stride; s.ntaps = (int)ph.taps.size();
    for (int t = 0; t < s.ntaps; ++t) { s.tdy[t] = (int8_t)ph.taps[t].dy; s.tdx[t] = (int8_t)ph.taps[t].dx; }
    s.w = ph.d_w_simt; s.Cin_pad = c.cin_pad; s.Cout = c.cout; s.Cout_pad8 = c.cout_pad8; s.bias = c.d_bias;
    s.Ho = c.transposed ? in.H : st.raw.H; s.Wo = c.transposed ? in.W : st.raw.W; s.raw = st.raw.p; s.raw_Cq = st.raw.Cq; s.raw_Wp = st.raw.Wp;
    s.raw_planar = st.raw.planar; s.raw_Hp = st.raw.Hp;
    s.oy_mul = s.ox_mul = c.out_mul; s.oy_off = ph.oy_off; s.ox_off = ph.ox_off;
    s.final_mode = last ? 1 : 0; s.out3 = out3; s.tanh_c = net->tanh_c;
    st.simt.push_back(s);
  }
  if (st.raw.planar) {  // residual-block kernel (conv_res.cu): cost-balanced tile table, same packed weights as conv_tc
    const ConvPhase &ph = c.phases[0];
    ResJob r;
    FAV_TRY(fill_res_job(c, ph, in, r));
    ResPlan rp = plan_res_tiles(r.Ho, r.Wo, net->num_sms);
    void *d_tiles = nullptr, *d_first = nullptr;
    FAV_TRY(check_cuda(cudaMalloc(&d_tiles, rp.tiles.size() * sizeof(ResTile)), "cudaMalloc(res tiles)"));
    pl.allocs.push_back(d_tiles);
    FAV_TRY(check_cuda(cudaMalloc(&d_first, rp.cta_first.size() * sizeof(int)), "cudaMalloc(res tiles)"));
    pl.allocs.push_back(d_first);
    FAV_TRY(check_cuda(cudaMemcpy(d_tiles, rp.tiles.data(), rp.tiles.size() * sizeof(ResTile), cudaMemcpyHostToDevice), "cudaMemcpy(res tiles)"));
    FAV_TRY(check_cuda(cudaMemcpy(d_first, rp.cta_first.data(), rp.cta_first.size() * sizeof(int), cudaMemcpyHostToDevice), "cudaMemcpy(res tiles)"));
    r.tiles = (const ResTile *)d_tiles; r.cta_first = (const int *)d_first; r.grid = rp.grid;
    r.b = ph.d_b_tc; r.bias = c.d_bias;
    r.raw = st.raw.p; r.raw_Hp = st.raw.Hp; r.raw_Wp = st.raw.Wp;
    st.res.push_back(r);
  }
  return FAV_OK;
}

static int build_plan(fav_net *net, int H, int W, Plan **out) {
  auto key = std::make_pair(H, W);
  auto it = net->plans.find(key);
  if (it != net->plans.end()) { *out = it->second.get(); net->last_plan = *out; return FAV_OK; }
  if (H % 4 || W % 4 || H < 16 || W < 16) {
    set_error("frame size %dx%d: H and W must be multiples of 4 (reflect-start nets restore the input size only "
              "then, SURVEY appendix B)", W, H);
    return FAV_ERR_INVALID;
  }
  std::unique_ptr<Plan> pl(new Plan());
  pl->H = H; pl->W = W;
  const int R = net->reflect_pad;
  if (R >= H || R >= W) { set_error("frame smaller than the reflection padding (%d)", R); return FAV_ERR_INVALID; }
  // pass 1: sizes -> largest raw tensor, stats slots
  size_t raw_max = 0;
  int stats_slots = 0;
  {
    int h = H + 2 * R, w = W + 2 * R;
    for (auto &op : net->ops) {
      if (op.kind == 2) { h *= op.scale; w *= op.scale; stats_slots += 2 * net->inorms[op.inorm[0]].C; continue; }
      int n = op.kind == 1 ? 2 : 1;
      for (int i = 0; i < n; ++i) {
        const ConvDef &c = net->convs[op.conv[i]];
        int ho, wo;
        conv_out_size(c, h, w, &ho, &wo);
        raw_max = std::max(raw_max, (size_t)(ho + 2) * round_up(c.cout, 4) * round_up(wo, kTileM));
        h = ho; w = wo;
        if (op.inorm[i] >= 0) stats_slots += 2 * c.cout;
      }
    }
  }
  FAV_TRY(alloc_zero(*pl, (void **)&pl->raw_buf, raw_max * sizeof(float) + 4096));
  FAV_TRY(alloc_zero(*pl, (void **)&pl->raw_buf2, raw_max * sizeof(float) + 4096));
  pl->stats_bytes = std::max(1, stats_slots) * sizeof(double);
  FAV_TRY(alloc_zero(*pl, (void **)&pl->stats, pl->stats_bytes));
  FAV_TRY(alloc_zero(*pl, (void **)&pl->msb, std::max(1, stats_slots) * 2 * sizeof(float)));
  FAV_TRY(alloc_zero(*pl, (void **)&pl->in7, (size_t)net->in_dim * H * W * sizeof(float)));
  FAV_TRY(alloc_zero(*pl, (void **)&pl->out3, (size_t)3 * H * W * sizeof(float)));

  // flat list of convs in execution order, to look up the consumer of each operand
  std::vector<int> order;
  for (auto &op : net->ops) {
    if (op.kind == 2) continue;
    order.push_back(op.conv[0]);
    if (op.kind == 1) order.push_back(op.conv[1]);
  }
  size_t pos = 0;
  int cur;
  FAV_TRY(make_operand(*pl, net->in_dim, H + 2 * R, W + 2 * R, &net->convs[order[0]], &cur));
  int stats_off = 0;
  for (size_t oi = 0; oi < net->ops.size(); ++oi) {
    const SpecOp &op = net->ops[oi];
    if (op.kind == 2) {  // UX: operand -> upsampled, normalised operand for the next convolution
      PlanStep us;
      us.kind = 2; us.scale = op.scale; us.inorm = op.inorm[0]; us.src = cur; us.relu = 1; us.stats_off = stats_off;
      const int C = net->inorms[op.inorm[0]].C;
      stats_off += 2 * C;
      const ConvDef *consumer = pos < order.size() ? &net->convs[order[pos]] : nullptr;
      FAV_TRY(make_operand(*pl, C, pl->ops[cur].H * op.scale, pl->ops[cur].W * op.scale, consumer, &us.dst));
      cur = us.dst;
      us.layer_index = (int)oi;
      pl->layer_operand[(int)oi] = cur;
      pl->steps.push_back(std::move(us));
      continue;
    }
    const int n = op.kind == 1 ? 2 : 1;
    const int block_in = cur;
    for (int i = 0; i < n; ++i, ++pos) {
      const ConvDef &c = net->convs[op.conv[i]];
      PlanStep cs;
      cs.kind = 0; cs.conv = op.conv[i]; cs.src = cur;
      int ho, wo;
      conv_out_size(c, pl->ops[cur].H, pl->ops[cur].W, &ho, &wo);
      cs.raw.p = (pos & 1) ? pl->raw_buf2 : pl->raw_buf; cs.raw.C = c.cout; cs.raw.Cq = round_up(c.cout, 4) / 4; cs.raw.H = ho; cs.raw.W = wo;
      cs.raw.Hp = ho + 2; cs.raw.Wp = round_up(wo, kTileM);
      const bool last = op.last && i == n - 1;
      if (!last && !getenv("FAV_NO_RES")) {  // residual-block convolutions: swapped-role kernel, planar raw output
        ConvDef &cm = net->convs[op.conv[i]];
        if (cm.phases.size() == 1 && conv_res_eligible(cm, cm.phases[0])) { cs.raw.planar = 1; cs.raw.Wp = round_up(wo, 16); }
      }
      FAV_TRY(build_conv_jobs(net, *pl, cs, c, last, nullptr));
      // second conv of a residual block: fold the preceding InstanceNorm + ReLU pass into its patch producers
      if (op.kind == 1 && i == 1 && !pl->steps.empty() && pl->steps.back().kind == 1 && !getenv("FAV_NO_NL")) {
        PlanStep &ap = pl->steps.back();
        const InDef &n = net->inorms[ap.inorm];
        const Operand &in = pl->ops[cs.src];
        bool ok = ap.skip < 0 && n.C <= 256 && cs.tc.size() == 1;
        if (ok && !cs.res.empty()) {  // conv_res.cu reads the PLANAR raw output of the block's first conv
          ok = ap.raw.planar && n.C == pl->ops[cs.src].C;
          if (ok) {
            ResJob &r = cs.res[0];
            const Operand &in = pl->ops[cs.src];
            r.nl = 1; r.nl_pad = c.pad; r.nl_raw = ap.raw.p; r.nl_Hp = ap.raw.Hp; r.nl_Wp = ap.raw.Wp; r.nl_H = ap.raw.H; r.nl_W = ap.raw.W;
            r.nl_relu = ap.relu; r.nl_C = n.C; r.nl_sums = pl->stats + ap.stats_off; r.nl_gamma = n.d_gamma; r.nl_beta = n.d_beta;
            r.nl_inv_count = 1.0 / ((double)ap.raw.H * ap.raw.W); r.nl_eps = 1e-5;
            (void)in;
            ap.fused_nl = true;
          }
        } else if (ok && ap.raw.planar) {
          ok = false;  // conv_tc's norm-on-load reads the quad layout only
        } else
        if (ok) {
          ConvJob &j = cs.tc[0];
          ok = !j.rf_R && !j.pf && !j.xfold_kw && j.nseg == 1 && j.seg_dst16[0] == 0 && j.seg_len16[0] == j.pslab16 && j.row_mul == 1;
          if (ok) {
            ConvJob t = j;
            t.nl = 1;
            if (const char *e = getenv("FAV_NL_MODE")) t.nl = atoi(e) == 2 ? 2 : 1;
            t.nl_raw = reinterpret_cast<const float4 *>(ap.raw.p); t.nl_Cq = ap.raw.Cq; t.nl_Wp = ap.raw.Wp;
            t.nl_H = ap.raw.H; t.nl_W = ap.raw.W; t.nl_padT = in.padT; t.nl_padL = in.padL; t.nl_relu = ap.relu; t.nl_C = n.C;
            t.nl_sums = pl->stats + ap.stats_off; t.nl_gamma = n.d_gamma; t.nl_beta = n.d_beta;
            t.nl_inv_count = 1.0 / ((double)ap.raw.H * ap.raw.W); t.nl_eps = 1e-5;
            conv_tc_choose_slots(t);
            if (conv_tc_smem_bytes(t) <= 227 * 1024 && (t.b_resident || t.b_slots >= 2)) { j = t; ap.fused_nl = true; }
          }
        }
      }
      if (last) { cs.layer_index = (int)oi; pl->steps.push_back(std::move(cs)); break; }
      RawTensor raw = cs.raw;
      pl->steps.push_back(std::move(cs));
      // InstanceNorm (+ReLU) (+skip) -> operand for the next convolution
      if (op.inorm[i] < 0 || c.cout % 8) {
        set_error("layer %s: a non-final convolution must be followed by InstanceNormalization (Cout %% 8 == 0)",
                  c.name.c_str());
        return FAV_ERR_UNSUPPORTED;
      }
      PlanStep ns;
      ns.kind = 1; ns.inorm = op.inorm[i]; ns.raw = raw; ns.stats_off = stats_off;
      for (ConvJob &j : pl->steps.back().tc) j.stats = pl->stats + stats_off;  // statistics fused into the epilogue
      for (ResJob &j : pl->steps.back().res) j.stats = pl->stats + stats_off;
      stats_off += 2 * c.cout;
      const ConvDef *consumer = pos + 1 < order.size() ? &net->convs[order[pos + 1]] : nullptr;
      FAV_TRY(make_operand(*pl, c.cout, ho, wo, consumer, &ns.dst));
      if (op.kind == 1) {
        ns.relu = i == 0 ? 1 : (op.relu ? 1 : 0);               // CX: needs_relu after the block (models_video.lua:108)
        ns.skip = (i == 1 && op.skip) ? block_in : -1;          // RX: ShaveImage(2) / Identity of the block input (:46-50)
        ns.shave = op.shave;
      } else {
        ns.relu = op.relu ? 1 : 0;
      }
      cur = ns.dst;
      if (i == n - 1) { ns.layer_index = (int)oi; pl->layer_operand[(int)oi] = cur; }
      pl->steps.push_back(std::move(ns));
    }
  }
  *out = pl.get();
  net->last_plan = pl.get();
  net->plans[key] = std::move(pl);
  return FAV_OK;
}

struct ProfRec {
  int kind;  // 0 pack, 1 conv, 2 in_stats(+finalize), 3 in_apply
  float ms;
  double work;  // conv: algorithmic FLOPs (logical channels); others: algorithmic bytes
  char name[24];
  cudaEvent_t e0, e1;
};

// in7 == nullptr: the caller has already written the first operand (fused temporal input, run_plan_frame)
static int run_plan(fav_net *net, Plan &pl, const float *in7, float *out3, int final_mode, cudaStream_t st,
                    std::vector<ProfRec> *prof = nullptr) {
  auto begin = [&](int kind, double work, const std::string &name) -> int {
    if (!prof) return FAV_OK;
    ProfRec r;
    r.kind = kind; r.ms = 0; r.work = work;
    snprintf(r.name, sizeof(r.name), "%s", name.c_str());
    FAV_TRY(check_cuda(cudaEventCreate(&r.e0), "cudaEventCreate"));
    FAV_TRY(check_cuda(cudaEventCreate(&r.e1), "cudaEventCreate"));
    FAV_TRY(check_cuda(cudaEventRecord(r.e0, st), "cudaEventRecord"));
    prof->push_back(r);
    return FAV_OK;
  };
  auto end = [&]() -> int {
    if (!prof) return FAV_OK;
    return check_cuda(cudaEventRecord(prof->back().e1, st), "cudaEventRecord");
  };
  FAV_TRY(check_cuda(cudaMemsetAsync(pl.stats, 0, pl.stats_bytes, st), "cudaMemsetAsync(stats)"));
  if (in7) {
    FAV_TRY(begin(0, (double)pl.ops[0].H * pl.ops[0].W * (4.0 * net->in_dim + 32.0), "pack_input"));
    FAV_TRY(launch_pack_input(in7, net->in_dim, pl.H, pl.W, net->reflect_pad, pl.ops[0], st));
    FAV_TRY(end());
  }
  for (PlanStep &s : pl.steps) {
    if (s.kind == 0) {
      {
        const ConvDef &c = net->convs[s.conv];
        const Operand &in = pl.ops[s.src];
        double px = c.transposed ? (double)in.H * in.W : (double)s.raw.H * s.raw.W;
        FAV_TRY(begin(1, 2.0 * c.cin * c.cout * c.k * c.k * px, c.name));
      }
      if (net->conv_impl == 0 && !s.res.empty()) {
        for (ResJob &j : s.res) FAV_TRY(launch_conv_res(j, st));
      } else if (net->conv_impl == 0) {
        for (ConvJob &j : s.tc) {
          if (j.final_mode) { j.final_mode = final_mode; j.out3 = out3; }
          FAV_TRY(launch_conv_tc(j, net->num_sms, st));
        }
      } else {
        for (SimtJob &j : s.simt) {
          if (j.final_mode) { j.final_mode = final_mode; j.out3 = out3; }
          FAV_TRY(launch_conv_simt(j, st));
        }
      }
      FAV_TRY(end());
    } else if (s.kind == 2) {
      const InDef &n = net->inorms[s.inorm];
      const Operand &src = pl.ops[s.src];
      FAV_TRY(begin(3, 12.0 * n.C * src.H * src.W * (1.0 + s.scale * s.scale) / 2.0, n.name + ".up"));
      FAV_TRY(launch_up_in(src, pl.stats + s.stats_off, n.d_gamma, n.d_beta, 1e-5f, s.relu, s.scale, pl.ops[s.dst], st));
      FAV_TRY(end());
    } else {
      if (s.fused_nl && net->conv_impl == 0) continue;  // the consumer conv normalises while it loads (norm-on-load)
      const InDef &n = net->inorms[s.inorm];
      double *sums = pl.stats + s.stats_off;
      const double elems = (double)n.C * s.raw.H * s.raw.W;
      if (net->conv_impl != 0) {  // the tcgen05 epilogue already accumulated the statistics
        FAV_TRY(begin(2, 4.0 * elems, n.name + ".stats"));
        FAV_TRY(launch_in_stats(s.raw, sums, st));
        FAV_TRY(end());
      }
      // InstanceNormalization.lua:21,39: eps = 1e-5, statistics over H*W of each (n, c)
      FAV_TRY(begin(3, (s.skip >= 0 ? 12.0 : 8.0) * elems, n.name + ".apply"));
      FAV_TRY(launch_in_apply(s.raw, sums, n.d_gamma, n.d_beta, 1e-5f, s.relu, s.skip >= 0 ? &pl.ops[s.skip] : nullptr, s.shave,
                              pl.ops[s.dst], st));
      FAV_TRY(end());
    }
  }
  return FAV_OK;
}

// run_[next_]image: the network part of a frame (input = pl.in7, fused deprocess) as one graph launch
static int run_plan_frame(fav_net *net, Plan &pl, const float *src7, float *out3, cudaStream_t st) {
  static const bool no_graph = getenv("FAV_NO_GRAPH") != nullptr;
  if (no_graph || net->conv_impl != 0) return run_plan(net, pl, src7, out3, 2, st);
  for (Plan::GraphEntry &g : pl.graphs)
    if (g.out3 == out3) {
      g.last_use = ++pl.use_clock;
      FAV_TRY(check_cuda(cudaGraphLaunch(g.exec, st), "cudaGraphLaunch"));
      g_launches.fetch_add(g.launches, std::memory_order_relaxed);
      return FAV_OK;
    }
  if (pl.eager_runs < 1) {  // one-time lazy initialisation (function attributes) must happen outside a capture
    ++pl.eager_runs;
    return run_plan(net, pl, src7, out3, 2, st);
  }
  if (!net->cap_stream)
    FAV_TRY(check_cuda(cudaStreamCreateWithFlags(&net->cap_stream, cudaStreamNonBlocking), "cudaStreamCreate(capture)"));
  const uint64_t l0 = g_launches.load();
  FAV_TRY(check_cuda(cudaStreamBeginCapture(net->cap_stream, cudaStreamCaptureModeRelaxed), "cudaStreamBeginCapture"));
  const int rc = run_plan(net, pl, src7, out3, 2, net->cap_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(net->cap_stream, &graph);
  const uint64_t n = g_launches.load() - l0;
  g_launches.fetch_sub(n, std::memory_order_relaxed);  // captured, not executed
  if (rc != FAV_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc != FAV_OK) return rc;
    return run_plan(net, pl, src7, out3, 2, st);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    cudaGetLastError();
    return run_plan(net, pl, src7, out3, 2, st);
  }
  if (pl.graphs.size() >= 8) {  // evict the least recently used destination
    size_t v = 0;
    for (size_t i = 1; i < pl.graphs.size(); ++i)
      if (pl.graphs[i].last_use < pl.graphs[v].last_use) v = i;
    cudaGraphExecDestroy(pl.graphs[v].exec);
    pl.graphs.erase(pl.graphs.begin() + v);
  }
  pl.graphs.push_back(Plan::GraphEntry{out3, exec, n, ++pl.use_clock});
  FAV_TRY(check_cuda(cudaGraphLaunch(exec, st), "cudaGraphLaunch"));
  g_launches.fetch_add(n, std::memory_order_relaxed);
  return FAV_OK;
}

}  // namespace fav

// =============================================================================================================
extern "C" {

int fav_net_create(const char *arch, const char *padding_type, float tanh_constant, int in_dim, fav_net_t **out) {
  FAV_REQUIRE(arch && out, "fav_net_create: null argument");
  FAV_REQUIRE(in_dim >= 1 && in_dim <= 8, "fav_net_create: in_dim must be in [1,8] (video nets use 7, image nets 3)");
  // models_video.lua:13-19,27-31,46-50,71-79: 'reflect-start' (every released video model; train_video.lua:25) = unpadded
  // residual convs + ShaveImage(2) + one lazily inserted reflection pad (train_video.lua:319-324); 'zero' = residual convs with
  // zero padding 1 and an Identity skip.  'reflect' / 'replicate' put a padding module in front of EVERY convolution and 'none'
  // returns a smaller frame than it was given: not built (no released model uses them).
  const std::string ptype = padding_type ? padding_type : "reflect-start";
  if (ptype != "reflect-start" && ptype != "zero") {
    set_error("padding_type '%s' is not supported (supported: 'reflect-start', the released video models' setting "
              "(train_video.lua:25), and 'zero')", ptype.c_str());
    return FAV_ERR_UNSUPPORTED;
  }
  const bool zero_pad = ptype == "zero";
  std::unique_ptr<fav_net> net(new fav_net());
  net->arch = arch; net->tanh_c = tanh_constant; net->in_dim = in_dim; net->padding_type = ptype;
  // tokenise (models_video.lua:56)
  std::vector<std::string> toks;
  {
    std::string s(arch), cur;
    for (char ch : s) {
      if (ch == ',') { toks.push_back(cur); cur.clear(); }
      else if (ch != ' ') cur.push_back(ch);
    }
    if (!cur.empty()) toks.push_back(cur);
  }
  FAV_REQUIRE(!toks.empty(), "fav_net_create: empty arch");
  int prev = in_dim;
  double scale = 1.0, shrink = 0.0;
  for (size_t i = 0; i < toks.size(); ++i) {
    const std::string &v = toks[i];
    const std::string name = "l" + std::to_string(i);
    SpecOp op;
    op.kind = 0; op.relu = true; op.last = (i + 1 == toks.size());
    bool needs_bn = true;
    int next = 0;
    char c0 = v[0];
    if (c0 == 'c' && v.size() >= 6 && v[2] == 's' && v[4] == '-') {  // cXsY-Z  models_video.lua:65-80
      int f = v[1] - '0', s = v[3] - '0';
      next = atoi(v.c_str() + 5);
      FAV_REQUIRE(f % 2 == 1 && (s == 1 || s == 2) && next > 0, "bad arch token '%s'", v.c_str());
      op.conv[0] = add_conv(net.get(), name, prev, next, f, s, (f - 1) / 2, false, 0);
      if (s == 2) scale *= 2;
    } else if (c0 == 'd') {  // dX  :90-93
      next = atoi(v.c_str() + 1);
      FAV_REQUIRE(next > 0, "bad arch token '%s'", v.c_str());
      op.conv[0] = add_conv(net.get(), name, prev, next, 3, 2, 1, false, 0);
      scale *= 2;
    } else if (c0 == 'u') {  // uX  :99-102  SpatialFullConvolution(3,3,2,2,1,1,1,1)
      next = atoi(v.c_str() + 1);
      FAV_REQUIRE(next > 0, "bad arch token '%s'", v.c_str());
      op.conv[0] = add_conv(net.get(), name, prev, next, 3, 2, 1, true, 1);
      scale /= 2;
    } else if (c0 == 'U') {  // UX  :94-98  SpatialUpSamplingNearest(X), followed by IN + ReLU (:121-130)
      int sc = atoi(v.c_str() + 1);
      FAV_REQUIRE(sc >= 1 && sc <= 4 && !op.last && prev % 8 == 0, "bad arch token '%s'", v.c_str());
      next = prev;
      op.kind = 2; op.scale = sc;
      op.inorm[0] = add_in(net.get(), name + ".n", next);
      needs_bn = false;
      scale /= sc;
    } else if (c0 == 'f' && v.size() >= 6 && v[2] == 's' && v[4] == '-') {  // fXsY-Z  :81-89  SpatialFullConvolution(f,f,s,s,p,p,s-1,s-1)
      int f = v[1] - '0', st = v[3] - '0';
      next = atoi(v.c_str() + 5);
      FAV_REQUIRE(f % 2 == 1 && (st == 1 || st == 2) && next > 0, "bad arch token '%s'", v.c_str());
      op.conv[0] = add_conv(net.get(), name, prev, next, f, st, (f - 1) / 2, true, st - 1);
      scale /= st;
    } else if (c0 == 'R' || c0 == 'C') {  // RX :109-114 residual block; CX :103-108 the same conv block without the skip
      next = atoi(v.c_str() + 1);
      FAV_REQUIRE(next == prev, "block %c%d needs %d input channels (got %d)", c0, next, next, prev);
      const int bp = zero_pad ? 1 : 0;  // build_conv_block :12-20
      op.kind = 1;
      op.skip = c0 == 'R';
      op.shave = zero_pad ? 0 : 2;      // ShaveImage(2) vs Identity :46-50
      op.conv[0] = add_conv(net.get(), name + ".c1", prev, next, 3, 1, bp, false, 0);
      op.inorm[0] = add_in(net.get(), name + ".n1", next);
      op.conv[1] = add_conv(net.get(), name + ".c2", next, next, 3, 1, bp, false, 0);
      op.inorm[1] = add_in(net.get(), name + ".n2", next);
      needs_bn = false; op.relu = c0 == 'C';  // :107-108, :113-114
      if (!zero_pad) shrink += 4 * scale;
    } else {
      set_error("arch token '%s' is not supported by the sm_100a path (supported: cXsY-Z, fXsY-Z, dX, uX, UX, CX, RX)", v.c_str());
      return FAV_ERR_UNSUPPORTED;
    }
    if (op.last) { needs_bn = false; op.relu = false; }  // :117-120
    if (needs_bn && op.kind == 0) op.inorm[0] = add_in(net.get(), name + ".n", next);
    net->ops.push_back(op);
    prev = next;
  }
  FAV_REQUIRE(prev == 3, "the last layer must produce 3 channels (got %d)", prev);
  FAV_REQUIRE(net->ops.back().kind == 0 && !net->convs[net->ops.back().conv[0]].transposed,
              "the last arch token must be a (non-transposed) convolution");
  FAV_REQUIRE(shrink == std::floor(shrink) && ((int)shrink) % 2 == 0, "unsupported shrink %f", shrink);
  net->reflect_pad = (int)shrink / 2;  // train_video.lua:319-324
  *out = net.release();
  return FAV_OK;
}

void fav_net_destroy(fav_net_t *net) { delete net; }

int fav_net_num_params(const fav_net_t *net) { return net ? (int)net->params.size() : 0; }

int fav_net_param_info(const fav_net_t *net, int index, char *name_out, int64_t shape_out[4], int64_t *numel) {
  FAV_REQUIRE(net && index >= 0 && index < (int)net->params.size(), "fav_net_param_info: bad index");
  const Param &p = net->params[index];
  if (name_out) { strncpy(name_out, p.name.c_str(), 63); name_out[63] = 0; }
  if (shape_out) for (int i = 0; i < 4; ++i) shape_out[i] = p.shape[i];
  if (numel) *numel = p.numel;
  return FAV_OK;
}

int fav_net_set_param(fav_net_t *net, const char *name, const float *host_data, int64_t numel) {
  FAV_REQUIRE(net && name && host_data, "fav_net_set_param: null argument");
  FAV_REQUIRE(!net->finalized, "fav_net_set_param: net already finalized");
  for (Param &p : net->params)
    if (p.name == name) {
      FAV_REQUIRE(numel == p.numel, "param %s: expected %lld elements, got %lld", name, (long long)p.numel,
                  (long long)numel);
      p.host.assign(host_data, host_data + numel);
      p.set = true;
      return FAV_OK;
    }
  set_error("fav_net_set_param: unknown parameter '%s'", name);
  return FAV_ERR_INVALID;
}

int fav_net_finalize(fav_net_t *net) {
  FAV_REQUIRE(net, "fav_net_finalize: null net");
  if (net->finalized) return FAV_OK;
  for (Param &p : net->params) FAV_REQUIRE(p.set, "fav_net_finalize: parameter %s was never set", p.name.c_str());
  FAV_TRY(require_device());
  cudaDeviceProp prop;
  FAV_TRY(check_cuda(cudaGetDevice(&net->device), "cudaGetDevice"));
  FAV_TRY(check_cuda(cudaGetDeviceProperties(&prop, net->device), "cudaGetDeviceProperties"));
  if (prop.major != 10) {
    set_error("libfav_b200 targets sm_100a (B200); device is sm_%d%d", prop.major, prop.minor);
    return FAV_ERR_UNSUPPORTED;
  }
  net->num_sms = prop.multiProcessorCount;
  for (ConvDef &c : net->convs) FAV_TRY(pack_conv_weights(net, c));
  for (InDef &n : net->inorms) {
    FAV_TRY(dev_upload(net, net->params[n.pw].host, &n.d_gamma));
    FAV_TRY(dev_upload(net, net->params[n.pb].host, &n.d_beta));
  }
  net->finalized = true;
  return FAV_OK;
}

int fav_net_set_conv_impl(fav_net_t *net, int impl) {
  FAV_REQUIRE(net && (impl == 0 || impl == 1), "fav_net_set_conv_impl: impl must be 0 (tcgen05) or 1 (CUDA cores)");
  net->conv_impl = impl;
  return FAV_OK;
}

int fav_net_forward(fav_net_t *net, const float *in7, int H, int W, float *out3, void *stream) {
  FAV_REQUIRE(net && in7 && out3, "fav_net_forward: null argument");
  FAV_REQUIRE(net->finalized, "fav_net_forward: call fav_net_finalize first");
  Plan *pl;
  FAV_TRY(build_plan(net, H, W, &pl));
  return run_plan(net, *pl, in7, out3, 1, (cudaStream_t)stream);
}

int fav_net_profile(fav_net_t *net, const float *in7, int H, int W, float *out3, int max_steps, int *kinds, float *ms,
                    double *work, char *names24, int *n_out, void *stream) {
  FAV_REQUIRE(net && in7 && out3 && kinds && ms && work && n_out, "fav_net_profile: null argument");
  FAV_REQUIRE(net->finalized, "fav_net_profile: call fav_net_finalize first");
  Plan *pl;
  FAV_TRY(build_plan(net, H, W, &pl));
  std::vector<ProfRec> prof;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = run_plan(net, *pl, in7, out3, 1, st, &prof);
  if (rc == FAV_OK) rc = check_cuda(cudaStreamSynchronize(st), "fav_net_profile sync");
  int n = 0;
  for (ProfRec &r : prof) {
    if (rc == FAV_OK && n < max_steps) {
      cudaEventElapsedTime(&r.ms, r.e0, r.e1);
      kinds[n] = r.kind; ms[n] = r.ms; work[n] = r.work;
      if (names24) memcpy(names24 + 24 * n, r.name, 24);
      ++n;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  *n_out = n;
  return rc;
}

int fav_net_layer_output(fav_net_t *net, int index, float *out, int *C, int *Hl, int *Wl, void *stream) {
  FAV_REQUIRE(net && net->last_plan, "fav_net_layer_output: run a forward first");
  Plan *pl = net->last_plan;
  auto it = pl->layer_operand.find(index);
  FAV_REQUIRE(it != pl->layer_operand.end(), "fav_net_layer_output: layer %d has no stored activation", index);
  const Operand &o = pl->ops[it->second];
  if (C) *C = o.C;
  if (Hl) *Hl = o.H;
  if (Wl) *Wl = o.W;
  if (!out) return FAV_OK;
  return launch_unpack_operand(o, out, (cudaStream_t)stream);
}

int fav_run_image(fav_net_t *net, const float *content, const float *fill, int H, int W, float *out_rgb,
                  void *stream) {
  FAV_REQUIRE(net && content && out_rgb, "fav_run_image: null argument");
  FAV_REQUIRE(net->finalized, "fav_run_image: call fav_net_finalize first");
  // model_img == nil: the video model on cat(pre(img), fill, zeros) (core.lua:133-138); a separate 3-channel image model
  // (-model_img, core.lua:61-68,146) gets pre(img) only -- the same fused input kernel, whose channels 3..7 are then zero
  FAV_REQUIRE(net->in_dim == 7 || net->in_dim == 3, "fav_run_image: video model (7 input channels) or image model (3) required");
  if (net->in_dim == 3) fill = nullptr;
  Plan *pl;
  FAV_TRY(build_plan(net, H, W, &pl));
  cudaStream_t st = (cudaStream_t)stream;
  static const bool fuse_pack = getenv("FAV_NO_FUSEPACK") == nullptr;
  if (fuse_pack) {
    FAV_TRY(launch_temporal_input_packed(content, nullptr, nullptr, nullptr, fill, nullptr, pl->ops[0], net->reflect_pad, H, W, 0, true, st));
    return run_plan_frame(net, *pl, nullptr, out_rgb, st);
  }
  FAV_TRY(launch_temporal_input(content, nullptr, nullptr, nullptr, fill, nullptr, pl->in7, H, W, 0, true, st));
  return run_plan_frame(net, *pl, pl->in7, out_rgb, st);  // deprocess fused into the last epilogue (core.lua:149)
}

int fav_run_next_image(fav_net_t *net, const float *content, const float *prev_rgb, const float *flow,
                       const float *cert, const float *fill, const float *flow_mask, int H, int W, int border_mode,
                       float *out_rgb, void *stream) {
  FAV_REQUIRE(net && content && prev_rgb && flow && cert && out_rgb, "fav_run_next_image: null argument");
  FAV_REQUIRE(net->finalized, "fav_run_next_image: call fav_net_finalize first");
  FAV_REQUIRE(net->in_dim == 7, "fav_run_next_image: video model (7 input channels) required");
  FAV_REQUIRE(border_mode == FAV_BORDER_PER_TAP || border_mode == FAV_BORDER_PAD_PIXEL, "bad border_mode");
  Plan *pl;
  FAV_TRY(build_plan(net, H, W, &pl));
  cudaStream_t st = (cudaStream_t)stream;
  static const bool fuse_pack = getenv("FAV_NO_FUSEPACK") == nullptr;
  if (fuse_pack) {
    FAV_TRY(launch_temporal_input_packed(content, prev_rgb, flow, cert, fill, flow_mask, pl->ops[0], net->reflect_pad, H, W,
                                         border_mode, false, st));
    return run_plan_frame(net, *pl, nullptr, out_rgb, st);
  }
  FAV_TRY(launch_temporal_input(content, prev_rgb, flow, cert, fill, flow_mask, pl->in7, H, W, border_mode, false, st));
  return run_plan_frame(net, *pl, pl->in7, out_rgb, st);  // core.lua:172-173
}

// a-8 with the certainty computed inside: the whole temporal stage (occlusion test or given certainty -> min filter -> warp
// -> preprocess -> mask -> concat) is ONE kernel that writes the network's first operand, followed by one graph launch.
int fav_run_next_image_flows(fav_net_t *net, const float *content, const float *prev_rgb, const float *flow_bw,
                             const float *flow_fw_uv, const float *cert_raw, const float *fill, const float *flow_mask,
                             int H, int W, int min_filter_r, int border_mode, float *out_rgb, void *stream) {
  FAV_REQUIRE(net && content && prev_rgb && flow_bw && out_rgb, "fav_run_next_image_flows: null argument");
  FAV_REQUIRE((flow_fw_uv != nullptr) != (cert_raw != nullptr), "fav_run_next_image_flows: give either the forward flow or a certainty plane");
  FAV_REQUIRE(net->finalized, "fav_run_next_image_flows: call fav_net_finalize first");
  FAV_REQUIRE(net->in_dim == 7, "fav_run_next_image_flows: video model (7 input channels) required");
  FAV_REQUIRE(border_mode == FAV_BORDER_PER_TAP || border_mode == FAV_BORDER_PAD_PIXEL, "bad border_mode");
  FAV_REQUIRE(min_filter_r >= 0 && (min_filter_r <= 1 || (min_filter_r & 1)) && min_filter_r <= 15, "occlusions_min_filter must be odd <= 15");
  Plan *pl;
  FAV_TRY(build_plan(net, H, W, &pl));
  cudaStream_t st = (cudaStream_t)stream;
  static const bool no_stage = getenv("FAV_NO_STAGE") != nullptr;  // A/B timing: the three separate kernels
  int rc = no_stage ? FAV_ERR_UNSUPPORTED
                    : launch_temporal_stage(content, prev_rgb, flow_bw, flow_fw_uv, cert_raw, fill, flow_mask, nullptr, nullptr,
                                            &pl->ops[0], net->reflect_pad, H, W, min_filter_r, border_mode, st);
  if (rc == FAV_ERR_UNSUPPORTED) {  // unaligned planes / W % 4 != 0: the three separate kernels
    const int64_t HW = (int64_t)H * W;
    if (!pl->cert_a) {
      FAV_TRY(alloc_zero(*pl, (void **)&pl->cert_a, HW * sizeof(float)));
      FAV_TRY(alloc_zero(*pl, (void **)&pl->cert_b, HW * sizeof(float)));
    }
    const float *cert = cert_raw;
    if (flow_fw_uv) {
      FAV_TRY(launch_consistency(flow_bw + HW, flow_bw, flow_fw_uv, flow_fw_uv + HW, nullptr, nullptr, 0.f, nullptr, pl->cert_a, W, H, st));
      cert = pl->cert_a;
    }
    if (min_filter_r > 1) {
      FAV_TRY(launch_min_filter(cert, pl->cert_b, 1, H, W, min_filter_r, st));
      cert = pl->cert_b;
    }
    rc = launch_temporal_input_packed(content, prev_rgb, flow_bw, cert, fill, flow_mask, pl->ops[0], net->reflect_pad, H, W,
                                      border_mode, false, st);
  }
  FAV_TRY(rc);
  return run_plan_frame(net, *pl, nullptr, out_rgb, st);
}
}
