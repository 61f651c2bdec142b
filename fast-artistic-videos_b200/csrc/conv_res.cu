// conv_res.cu -- tcgen05 kernel of the residual-block convolutions (3x3, stride 1, 128 output channels) with the GEMM roles
// swapped: weights = M operand (128 TMEM lanes = output channels), pixels = N operand (TMEM columns).  See conv_res.cuh for
// the why; models_video.lua:20,32,41-53 for the layers; InstanceNormalization.lua:33-53 for the fused statistics.
//
// Warp roles (512 threads = 16 warps, 1 CTA / SM, persistent over the CTA's slice of the tile table):
//   warps 0-3    epilogue group 0: TMEM lane quarter = warp % 4, thread = one output channel; tcgen05.ld 16 pixels ->
//                +bias -> two 256-bit stores into the planar raw tensor; per-thread sum / sum of squares (InstanceNorm)
//   warps 4-7    operand input: epilogue group 1 (the groups take one output row of the pair each)
//                norm-on-load input: patch producers
//   warps 8-11   norm-on-load input: patch producers (raw fp32 of the previous conv -> InstanceNorm + ReLU -> fp16 hi/lo,
//                written MMA-ready into the stage); idle otherwise
//   warp  12     operand input: patch producer (16 bulk copies per stage, one per lane)
//   warp  13     weight producer (ring of 24 KB chunks) + TMEM allocation
//   warps 14,15  MMA issuers, one output row of the pair each.  They are the HIGHEST warp ids of their scheduler
//                partitions (14 % 4 = 2, 15 % 4 = 3): the warp arbiter prefers high warp ids (B300_MICROARCH.md), and an
//                issuer that loses issue slots to an ALU-heavy producer / epilogue warp starves the tensor pipe.  (In
//                conv_tc.cu the issuers are warps 6/7 below the epilogue / producer warps 8-14; its norm-on-load launches
//                issued a 2-row unit in 36k cycles against 31k without the extra producer warps -- tools/trace_conv.py timelines of the first
//                design, round 2.)
// Pipelines: 4 patch stages (4 rows x 2 channel blocks x (nt + 2) pixels, hi + lo), 3 weight slots, 2 TMEM accumulator
// stages (2 rows x 128 columns each) so that the epilogue of tile i overlaps the MMAs of tile i + 1.
#include <atomic>

#include "conv_res.cuh"
#include "tc_common.cuh"

namespace fav {

constexpr int kResThreads = 512;
constexpr int kResStages = 4;
constexpr int kResSlots = 3;
constexpr int kResNlWarps = 8;   // warps 4..11
constexpr int kResNlPx = 5;      // pixels per lane and slab: 5 x 32 >= 130
constexpr int kResNlMaxC = 256;
constexpr uint32_t kResStageBytes = 4u * kResCbG * kResPslab * 16u;  // one plane (hi or lo) of a stage: 33280 B
constexpr uint32_t kResChunkBytes = 2u * kResSpc * 2u * 128u * 16u;  // [hi|lo][3 steps][2 k-halves][128 couts] x 16 B = 24576 B

struct __align__(16) ResShared {
  uint64_t a_full[kResStages], a_empty[kResStages], b_full[kResSlots], b_empty[kResSlots], t_full[2], t_empty[2];
  uint32_t tmem_base, pad_;
  float nl_tab[3 * kResNlMaxC];  // mean | gamma * rstd | beta of the input channels (norm-on-load)
  float stat_x[2 * 128];         // partial statistics of epilogue group 1, added by group 0
};

size_t conv_res_smem_bytes(int) { return (size_t)kResStages * 2 * kResStageBytes + (size_t)kResSlots * kResChunkBytes + sizeof(ResShared) + 128; }
int conv_res_slots() { return kResSlots; }

__device__ __forceinline__ void st_global_v8(float *p, const float (&v)[16], int o) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[o]), "f"(v[o + 1]), "f"(v[o + 2]), "f"(v[o + 3]),
               "f"(v[o + 4]), "f"(v[o + 5]), "f"(v[o + 6]), "f"(v[o + 7])
               : "memory");
}

__device__ __forceinline__ void res_trace(const ResJob &job, int word, long long v) {
  if (job.trace) job.trace[(size_t)blockIdx.x * kTraceWords + word] = (unsigned long long)v;
}

// One (patch row, channel block) slab of a norm-on-load stage: raw fp32 planes -> InstanceNorm (+ReLU) -> fp16 hi/lo, stored
// MMA-ready at drow (hi) / drow + kResStageBytes (lo); lanes run along x, all loads of the slab (up to 40 independent
// 128-byte-coalesced requests per lane) are requested before the first is used.
// (Tried and removed: also adding the previous block's skip here and writing the block input back, which would delete the
// in_apply pass BETWEEN residual blocks -- with 16 more registers of skip data per pixel only 2 pixels per lane fit in flight,
// the slab needs two latency-exposed passes and the conv became producer bound: 70 us against 45.5 + 23 us. DESIGN.md 9.)
__device__ __forceinline__ void nl_slab(const ResJob &job, const float *nl_tab, uint8_t *drow, int cb, int ry, bool row_ok, int xb,
                                        int npx, int lane, int64_t pstride) {
  const float *p0 = job.nl_raw + ((int64_t)(cb * 8) * job.nl_Hp + (row_ok ? ry : 0)) * job.nl_Wp;
  float tm[8], ts[8], tb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { tm[i] = nl_tab[cb * 8 + i]; ts[i] = nl_tab[kResNlMaxC + cb * 8 + i]; tb[i] = nl_tab[2 * kResNlMaxC + cb * 8 + i]; }
  float v[kResNlPx][8];
#pragma unroll
  for (int k = 0; k < kResNlPx; ++k)
    if (k * 32 < npx) {
      const int x = xb + k * 32 + lane;
      const int xc = x < 0 ? 0 : (x < job.nl_W ? x : job.nl_W - 1);  // clamped (always valid) addresses
#pragma unroll
      for (int i = 0; i < 8; ++i) v[k][i] = __ldg(p0 + i * pstride + xc);
    }
#pragma unroll
  for (int k = 0; k < kResNlPx; ++k) {
    const int p = k * 32 + lane, x = xb + p;
    if (p < npx) {
      const bool ok = row_ok && x >= 0 && x < job.nl_W;  // outside the image: zero (never normalised)
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = (v[k][2 * i] - tm[2 * i]) * ts[2 * i] + tb[2 * i], b = (v[k][2 * i + 1] - tm[2 * i + 1]) * ts[2 * i + 1] + tb[2 * i + 1];
        if (job.nl_relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
        if (!ok) { a = 0.f; b = 0.f; }
        split_pair(a, b, h[i], l[i]);  // same values as split_store8
      }
      *reinterpret_cast<uint4 *>(drow + (uint32_t)p * 16u) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4 *>(drow + kResStageBytes + (uint32_t)p * 16u) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

__global__ void __launch_bounds__(kResThreads, 1) conv_res_kernel(const __grid_constant__ ResJob job) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t *a_base = smem;                                           // stage s: [hi plane][lo plane]
  uint8_t *b_base = a_base + kResStages * 2 * kResStageBytes;       // slot s: [hi: 3 steps][lo: 3 steps]
  ResShared *sh = reinterpret_cast<ResShared *>(b_base + kResSlots * kResChunkBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool nl = job.nl != 0;
  if (job.trace && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    res_trace(job, 0, (long long)gt);
    res_trace(job, 1, clock64());
  }
  const int t_begin = __ldg(job.cta_first + blockIdx.x), t_end = __ldg(job.cta_first + blockIdx.x + 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kResStages; ++i) { mbar_init(&sh->a_full[i], nl ? kResNlWarps : 1); mbar_init(&sh->a_empty[i], 2); }
    for (int i = 0; i < kResSlots; ++i) { mbar_init(&sh->b_full[i], 1); mbar_init(&sh->b_empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh->t_full[i], 2); mbar_init(&sh->t_empty[i], nl ? 128 : 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  if (threadIdx.x == 0 && job.trace) res_trace(job, 2, clock64());
  const int ngroups = job.ngroups;

  if (nl && warp >= 4 && warp < 12) {
    // ===== norm-on-load patch producers =====
    // The 8 producer warps first finalise mean / gamma * rstd / beta of the input channels from the producer conv's fused
    // statistics (same arithmetic as in_apply_kernel: biased variance, eps inside the sqrt, InstanceNormalization.lua:39-50)
    // and meet at a named barrier of their own; weight producer and MMA warps are already running.
    {
      const int c = (int)threadIdx.x - 128;
      if (c < job.nl_C) {
        in_finalize(job.nl_sums[c], job.nl_sums[job.nl_C + c], job.nl_inv_count, job.nl_eps, job.nl_gamma[c], sh->nl_tab[c],
                    sh->nl_tab[kResNlMaxC + c]);
        sh->nl_tab[2 * kResNlMaxC + c] = job.nl_beta[c];
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
    }
    const int pw = warp - 4;
    const int64_t pstride = (int64_t)job.nl_Hp * job.nl_Wp;  // between channel planes
    uint32_t s = 0, ph = 0;
    for (int ti = t_begin; ti < t_end; ++ti) {
      const ResTile t = job.tiles[ti];
      const int npx = t.nt + 2, xb = t.x0 - job.nl_pad;
      for (int g = 0; g < ngroups; ++g) {
        mbar_wait(&sh->a_empty[s], ph ^ 1);
        uint8_t *stage = a_base + s * 2 * kResStageBytes;
        for (int rc = pw; rc < 4 * kResCbG; rc += kResNlWarps) {  // one (patch row, channel block) slab per warp and stage
          const int ri = rc / kResCbG, cb = g * kResCbG + (rc % kResCbG);
          const int ry = t.y + ri - job.nl_pad;
          const bool row_ok = ry >= 0 && ry < job.nl_H;
          uint8_t *drow = stage + (uint32_t)(rc * kResPslab) * 16u;
          nl_slab(job, sh->nl_tab, drow, cb, ry, row_ok, xb, npx, lane, pstride);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->a_full[s]);
        if (++s == kResStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (!nl && warp == 12) {
    // ===== patch producer (operand input): lane = (patch row, channel block, hi/lo) =====
    constexpr int kCopies = 4 * kResCbG * 2;  // per stage: 4 patch rows x 2 channel blocks x (hi, lo) = 16 <= 32 lanes
    const int ri = lane / (2 * kResCbG), cbi = (lane >> 1) % kResCbG, part = lane & 1;
    const uint4 *plane = part ? job.a_lo : job.a_hi;
    const uint32_t dst_off = (uint32_t)part * kResStageBytes + (uint32_t)((ri * kResCbG + cbi) * kResPslab) * 16u;
    uint32_t s = 0, ph = 0;
    for (int ti = t_begin; ti < t_end; ++ti) {
      const ResTile t = job.tiles[ti];
      const uint32_t seg_bytes = (uint32_t)(t.nt + 2) * 16u;
      for (int g = 0; g < ngroups; ++g) {
        mbar_wait(&sh->a_empty[s], ph ^ 1);
        // start of the kernel: all 148 CTAs fill their rings at once (30 MB burst) and the first MMA needs only stage 0 +
        // weight slot 0 -- hold the copies of stages 2.. back until stage 0 has landed (one stage stays in flight behind it)
        if (ti == t_begin && g == 2) mbar_wait(&sh->a_full[0], 0);
        if (lane == 0) mbar_arrive_expect_tx(&sh->a_full[s], (uint32_t)kCopies * seg_bytes);
        __syncwarp();
        if (lane < kCopies) {
          const int64_t src16 = ((int64_t)(t.y + ri + job.in_row0) * job.a_Cb + g * kResCbG + cbi) * job.a_slab16 + t.x0 + job.in_col0;
          bulk_g2s(a_base + s * 2 * kResStageBytes + dst_off, plane + src16, seg_bytes, &sh->a_full[s]);
        }
        if (++s == kResStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 13) {
    // ===== weight producer: 6 chunks per channel group through a ring of kResSlots slots =====
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int ti = t_begin; ti < t_end; ++ti)
        for (int gc = 0; gc < ngroups * kResChunks; ++gc) {
          mbar_wait(&sh->b_empty[s], ph ^ 1);
          if (ti == t_begin && gc == 2) mbar_wait(&sh->b_full[0], 0);  // same: slot 2 waits until slot 0 has landed
          mbar_arrive_expect_tx(&sh->b_full[s], kResChunkBytes);
          bulk_g2s(b_base + s * kResChunkBytes, job.b + (int64_t)gc * (kResChunkBytes / 16), kResChunkBytes, &sh->b_full[s]);
          if (++s == kResSlots) { s = 0; ph ^= 1; }
        }
    }
    __syncwarp();
  } else if (warp >= 14) {
    // ===== MMA issuers: warp 14 = first output row of the pair, warp 15 = second (one patch row lower).  The whole warp
    // runs the warp-uniform control flow (descriptors in uniform registers), one elected lane issues =====
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t r = (uint32_t)(warp - 14);
    // matrix descriptors (K-major, no swizzle): high word = SBO 8 (128 B between 8-row groups) | version 1 at bit 46
    const uint32_t desc_hi = 8u | (1u << 14);
    constexpr uint32_t w_step16 = 2u * 128u, w_lo16 = kResSpc * w_step16;  // one K16 step of weights; lo image after the hi image
    constexpr uint32_t p_row16 = kResCbG * kResPslab, p_lo16 = kResStageBytes / 16u;
    uint32_t sa = 0, aph = 0, sb = 0, bph = 0, tl = 0;
    for (int ti = t_begin; ti < t_end; ++ti, ++tl) {
      const ResTile t = job.tiles[ti];
      const uint32_t as = tl & 1, tph = (tl >> 1) & 1;
      mbar_wait(&sh->t_empty[as], tph ^ 1);
      tc_fence_after();
      const bool tr = job.trace && warp == 14 && lane == 0 && tl < (uint32_t)kTraceUnits;
      long long tr_a = 0, tr_b = 0;
      if (tr) res_trace(job, 8 + 8 * (int)tl, clock64());
      // instruction descriptor: D = f32, A = B = f16, K-major both, N = nt (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
      const uint32_t idesc = (1u << 4) | ((uint32_t)(t.nt >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t d = tmem_base + as * 256u + r * 128u;
      uint32_t accumulate = 0;
      for (int g = 0; g < ngroups; ++g) {
        const long long tw = tr ? clock64() : 0;
        mbar_wait(&sh->a_full[sa], aph);
        tc_fence_after();
        if (tr) { const long long now = clock64(); tr_a += now - tw; if (g == 0) res_trace(job, 8 + 8 * (int)tl + 1, now); }
        const uint32_t p_hi16 = (smem_u32(a_base + sa * 2 * kResStageBytes) >> 4) + r * p_row16, p_lo = p_hi16 + p_lo16;
        int sidx = 0;
        for (int c = 0; c < kResChunks; ++c) {
          const long long tw2 = tr ? clock64() : 0;
          mbar_wait(&sh->b_full[sb], bph);
          tc_fence_after();
          if (tr) tr_b += clock64() - tw2;
          const uint32_t w16 = (smem_u32(b_base + sb * kResChunkBytes) >> 4) | (128u << 16);  // LBO = 128 rows x 16 B
#pragma unroll
          for (int st = 0; st < kResSpc; ++st, ++sidx) {
            const uint32_t dls = job.steps[sidx];
            const uint32_t ws = w16 + (uint32_t)st * w_step16;
            const uint64_t pd_hi = ((uint64_t)desc_hi << 32) | (p_hi16 + dls), pd_lo = ((uint64_t)desc_hi << 32) | (p_lo + dls);
            const uint64_t wd_hi = ((uint64_t)desc_hi << 32) | ws, wd_lo = ((uint64_t)desc_hi << 32) | (ws + w_lo16);
            if (leader) {
              tc_mma_f16(d, wd_hi, pd_hi, idesc, accumulate);  // hi * hi
              tc_mma_f16(d, wd_lo, pd_hi, idesc, 1);           // lo(w) * hi(x)
              tc_mma_f16(d, wd_hi, pd_lo, idesc, 1);           // hi(w) * lo(x)
            }
            accumulate = 1;
          }
          if (leader) tc_commit(&sh->b_empty[sb]);  // frees the weight slot when this warp's MMAs retire
          if (++sb == kResSlots) { sb = 0; bph ^= 1; }
        }
        if (leader) tc_commit(&sh->a_empty[sa]);    // frees the patch stage
        if (++sa == kResStages) { sa = 0; aph ^= 1; }
      }
      if (leader) tc_commit(&sh->t_full[as]);       // this row's accumulator is complete
      if (tr) { res_trace(job, 8 + 8 * (int)tl + 2, clock64()); res_trace(job, 8 + 8 * (int)tl + 3, tr_a); res_trace(job, 8 + 8 * (int)tl + 4, tr_b); }
    }
    __syncwarp();
  } else if (warp < 4 || (!nl && warp < 8)) {
    // ===== epilogue: thread = output channel (TMEM lane), 16 consecutive pixels per tcgen05.ld =====
    const int q = warp & 3, eg = warp >> 2, neg = nl ? 1 : 2;
    const int cout = q * 32 + lane;
    const float bias = __ldg(job.bias + cout);
    float ssum = 0.f, qsum = 0.f;
    float *plane = job.raw + (int64_t)cout * job.raw_Hp * job.raw_Wp;
    uint32_t tl = 0;
    for (int ti = t_begin; ti < t_end; ++ti, ++tl) {
      const ResTile t = job.tiles[ti];
      const uint32_t as = tl & 1, tph = (tl >> 1) & 1;
      mbar_wait(&sh->t_full[as], tph);
      tc_fence_after();
      const bool etr = job.trace && threadIdx.x == 0 && tl < (uint32_t)kTraceUnits;
      if (etr) res_trace(job, 8 + 8 * (int)tl + 5, clock64());
      for (int row = eg; row < 2; row += neg) {
        const int yo = t.y + row;
        if (yo >= job.Ho) continue;  // odd Ho: the second row of the last pair does not exist
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256u + (uint32_t)row * 128u;
        float *out = plane + (int64_t)yo * job.raw_Wp + t.x0;
        for (int c0 = 0; c0 < t.nt; c0 += 16) {
          uint32_t rr[16];
          tmem_ld16(taddr + (uint32_t)c0, rr);
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]) + bias;
          st_global_v8(out + c0, v, 0);  // raw_Wp >= round_up(Wo, 16): the columns beyond Wo are never read
          st_global_v8(out + c0 + 8, v, 8);
          const int nvalid = job.Wo - (t.x0 + c0);
          if (nvalid >= 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { ssum += v[i]; qsum = fmaf(v[i], v[i], qsum); }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nvalid) { ssum += v[i]; qsum = fmaf(v[i], v[i], qsum); }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&sh->t_empty[as]);  // TMEM stage drained
      if (etr) res_trace(job, 8 + 8 * (int)tl + 6, clock64());
    }
    if (job.stats) {
      // two epilogue groups hold partial sums of the same channels (one output row each): group 1 -> shared memory ->
      // group 0, fixed order; then one double atomic per channel and quantity per CTA (static tile -> CTA assignment)
      if (neg == 2) {
        if (eg == 1) { sh->stat_x[cout] = ssum; sh->stat_x[128 + cout] = qsum; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (eg == 0) { ssum += sh->stat_x[cout]; qsum += sh->stat_x[128 + cout]; }
      }
      if (eg == 0) {
        atomicAdd(job.stats + cout, (double)ssum);
        atomicAdd(job.stats + 128 + cout, (double)qsum);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && job.trace) { res_trace(job, 3, clock64()); res_trace(job, 4, t_end - t_begin); }
  if (warp == 13) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// timeline buffer shared with conv_tc.cu's launches (fav_debug_set_trace)
unsigned long long *conv_trace_claim(size_t words);

int launch_conv_res(const ResJob &job_in, cudaStream_t st) {
  ResJob job = job_in;
  job.trace = conv_trace_claim((size_t)job.grid * kTraceWords);
  const size_t smem = conv_res_smem_bytes(kResSlots);
  static std::atomic<uint64_t> attr_set{0};  // per-device function attribute
  int dev = 0;
  FAV_TRY(check_cuda(cudaGetDevice(&dev), "cudaGetDevice"));
  const uint64_t bit = 1ull << (dev & 63);
  if (!(attr_set.load(std::memory_order_acquire) & bit)) {
    FAV_TRY(check_cuda(cudaFuncSetAttribute(conv_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute(conv_res)"));
    attr_set.fetch_or(bit, std::memory_order_release);
  }
  conv_res_kernel<<<job.grid, kResThreads, smem, st>>>(job);
  return post_launch("conv_res");
}

}  // namespace fav
