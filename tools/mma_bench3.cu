// mma_bench3.cu -- replicate conv_tc_kernel's issue loop for the residual convs (N = 128, two accumulators per unit,
// 4 channel groups x 9 weight chunks x 2 K steps x 6 MMAs) in isolation and bisect the per-step overhead.  Timing only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
struct Cfg {
  int units, groups, chunks, spc, two;
  int N;      // 0 = 128
  int flags;  // 1 wait+fence per chunk and per group, 2 commit per chunk, 4 per-chunk divergent region (else per step), 8 no LDC table
              // 16 one MMA of three
  uint32_t steps[32];
  uint64_t steps64[32];
  uint32_t desc_hi_opaque;
};
__global__ void __launch_bounds__(224, 1) bench(const __grid_constant__ Cfg c, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done_bar, ready_bar, sink[8];
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u ^ ((uint32_t)i * 2654435761u & 0x03ff03ffu);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar)), "r"((c.flags & 64) ? 2 : 1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ready_bar)));
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sink[i])));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const bool dual = c.flags & 64;
  const int iw = threadIdx.x >> 5;
  if (threadIdx.x < 32 || (dual && iw == 1)) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t N = c.N ? (uint32_t)c.N : 128u;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t desc_hi = 8u | (1u << 14);
    const uint32_t a_hi16 = smem_u32(smem) >> 4, a_lo16 = a_hi16 + (40960 >> 4), a_tile16 = 4 * 130;
    const uint32_t b_step16 = 2u * N, b_lo16 = (uint32_t)c.spc * b_step16;
    const bool two = c.two && !dual, one = c.flags & 16;
    const uint32_t drow = dual ? (uint32_t)iw : 0u;  // dual issue: warp 1 owns the second accumulator row
    const int spc = c.spc;
    long long t0 = clock64();
    uint32_t sb = 0;
    for (int u = 0; u < c.units; ++u) {
      const uint32_t d0 = tmem + (u & 1) * 256u + drow * 128u, d1 = d0 + 128u;
      uint32_t acc = 0, sc = 0;
      for (int g = 0; g < c.groups; ++g) {
        if (c.flags & 1) { wait(&ready_bar, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        int sidx = 0;
        for (int ch = 0; ch < c.chunks; ++ch) {
          if (c.flags & 1) { wait(&ready_bar, 0); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
          const uint32_t bh = ((smem_u32(smem + 100 * 1024) >> 4) + sb * 1024u) | (N << 16);
          if (c.flags & 256) {
            // the generic issue loop of conv_tc.cu as of round 2: plain loop in one divergent region per chunk; K-split (flag 128)
            if (leader) {
              uint32_t bs = bh;
              for (uint32_t st = 0; st < (uint32_t)spc; ++st, bs += b_step16) {
                if ((c.flags & 128) && ((sc + st) & 1u) != drow) continue;
                const uint32_t dls = c.steps[sidx + st];
                const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | (a_hi16 + dls), ad_lo = ((uint64_t)desc_hi << 32) | (a_lo16 + dls);
                const uint64_t bd_hi = ((uint64_t)desc_hi << 32) | bs, bd_lo = ((uint64_t)desc_hi << 32) | (bs + b_lo16);
                mma(d0, ad_hi, bd_hi, idesc, acc);
                if (!one) { mma(d0, ad_lo, bd_hi, idesc, 1); mma(d0, ad_hi, bd_lo, idesc, 1); }
                acc = 1;
              }
            }
            sidx += spc; sc += (uint32_t)spc;
          } else if (c.flags & 32) {
            // slim descriptors: 64-bit invariant bases + one 64-bit add per descriptor
            const uint64_t A_HI = ((uint64_t)c.desc_hi_opaque << 32) | (a_hi16 + drow * a_tile16), A_LO = ((uint64_t)c.desc_hi_opaque << 32) | (a_lo16 + drow * a_tile16);
            const uint64_t BH = ((uint64_t)c.desc_hi_opaque << 32) | bh;
            if (leader) {
#pragma unroll
              for (int st = 0; st < 4; ++st)
                if (st < spc) {
                  const uint64_t dls = c.steps64[sidx + st];
                  const uint64_t ad_hi = A_HI + dls, ad_lo = A_LO + dls;
                  const uint64_t bd_hi = BH + (uint64_t)(st * 256), bd_lo = bd_hi + b_lo16;
                  mma(d0, ad_hi, bd_hi, idesc, acc);
                  if (!one) { mma(d0, ad_lo, bd_hi, idesc, 1); mma(d0, ad_hi, bd_lo, idesc, 1); }
                  if (two) {
                    mma(d1, ad_hi + a_tile16, bd_hi, idesc, acc);
                    if (!one) { mma(d1, ad_lo + a_tile16, bd_hi, idesc, 1); mma(d1, ad_hi + a_tile16, bd_lo, idesc, 1); }
                  }
                  acc = 1;
                }
            }
            sidx += spc;
          } else if (c.flags & 4) {
            if (leader) {
#pragma unroll
              for (int st = 0; st < 4; ++st)
                if (st < spc) {
                  const uint32_t dls = (c.flags & 8) ? ((uint32_t)(sidx + st) * 260u | (130u << 16)) : c.steps[sidx + st];
                  const uint32_t bs = bh + (uint32_t)st * b_step16;
                  const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | (a_hi16 + dls), ad_lo = ((uint64_t)desc_hi << 32) | (a_lo16 + dls);
                  const uint64_t bd_hi = ((uint64_t)desc_hi << 32) | bs, bd_lo = ((uint64_t)desc_hi << 32) | (bs + b_lo16);
                  mma(d0, ad_hi, bd_hi, idesc, acc);
                  if (!one) { mma(d0, ad_lo, bd_hi, idesc, 1); mma(d0, ad_hi, bd_lo, idesc, 1); }
                  if (two) {
                    mma(d1, ad_hi + a_tile16, bd_hi, idesc, acc);
                    if (!one) { mma(d1, ad_lo + a_tile16, bd_hi, idesc, 1); mma(d1, ad_hi + a_tile16, bd_lo, idesc, 1); }
                  }
                  acc = 1;
                }
            }
            sidx += spc;
          } else {
            for (int st = 0; st < spc; ++st, ++sidx) {
              const uint32_t dls = (c.flags & 8) ? ((uint32_t)sidx * 260u | (130u << 16)) : c.steps[sidx];
              const uint32_t bs = bh + (uint32_t)st * b_step16;
              const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | (a_hi16 + dls), ad_lo = ((uint64_t)desc_hi << 32) | (a_lo16 + dls);
              const uint64_t bd_hi = ((uint64_t)desc_hi << 32) | bs, bd_lo = ((uint64_t)desc_hi << 32) | (bs + b_lo16);
              if (leader) {
                mma(d0, ad_hi, bd_hi, idesc, acc);
                if (!one) { mma(d0, ad_lo, bd_hi, idesc, 1); mma(d0, ad_hi, bd_lo, idesc, 1); }
                if (two) {
                  mma(d1, ad_hi + a_tile16, bd_hi, idesc, acc);
                  if (!one) { mma(d1, ad_lo + a_tile16, bd_hi, idesc, 1); mma(d1, ad_hi + a_tile16, bd_lo, idesc, 1); }
                }
              }
              acc = 1;
            }
          }
          acc = 1;
          if ((c.flags & 2) && leader) commit(&sink[sb & 7]);
          if (++sb == 5) sb = 0;
        }
        if ((c.flags & 2) && leader) commit(&sink[7]);
      }
      if (leader) commit(&sink[6]);
    }
    if (leader) commit(&done_bar);
    wait(&done_bar, 0);
    long long t1 = clock64();
    if (leader && iw == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long *d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char *name; int units, groups, chunks, spc, two, flags, N; } cfgs[] = {
      {"res: per-step region, waits, commits (kernel)", 4, 4, 9, 2, 1, 3},
      {"res: per-step region, no waits",                4, 4, 9, 2, 1, 2},
      {"res: per-step region, no waits/commits",        4, 4, 9, 2, 1, 0},
      {"res: per-step region, computed steps",          4, 4, 9, 2, 1, 3 | 8},
      {"res: per-chunk region, waits, commits",         4, 4, 9, 2, 1, 3 | 4},
      {"res: per-chunk region, computed steps",         4, 4, 9, 2, 1, 3 | 4 | 8},
      {"res: per-step, one MMA of three",               4, 4, 9, 2, 1, 3 | 16},
      {"res: single accumulator (mt=1)",                8, 4, 9, 2, 0, 3},
      {"res: spc 4 chunks",                             4, 4, 9, 4, 1, 3 | 4},
      {"res: one of three, no waits/commits",           4, 4, 9, 2, 1, 16},
      {"res: one of three, waits only",                 4, 4, 9, 2, 1, 16 | 1},
      {"res: one of three, commits only",               4, 4, 9, 2, 1, 16 | 2},
      {"res: slim descriptors, waits+commits",          4, 4, 9, 2, 1, 3 | 32},
      {"res: slim, one of three",                       4, 4, 9, 2, 1, 3 | 32 | 16},
      {"res: slim, no waits/commits",                   4, 4, 9, 2, 1, 32},
      {"res: dual issue warps, per-step, waits+commits",4, 4, 9, 2, 1, 3 | 64},
      {"res: dual issue warps, slim",                   4, 4, 9, 2, 1, 3 | 32 | 64},
      {"res: dual issue warps, slim, one of three",     4, 4, 9, 2, 1, 3 | 32 | 64 | 16},
      {"mt1: slim",                                     8, 4, 9, 2, 0, 3 | 32},
      // r02: the narrow layers of conv_tc (generic loop as of round 2 = flag 256; 128 = K-split; 64 = two issuing warps)
      {"d64 : lean, ksplit, waits+commits",             17, 2, 3, 3, 0, 3 | 64 | 128 | 256, 64},
      {"d64 : lean, ksplit, no waits",                  17, 2, 3, 3, 0, 2 | 64 | 128 | 256, 64},
      {"d64 : lean, ksplit, no waits/commits",          17, 2, 3, 3, 0, 64 | 128 | 256, 64},
      {"d64 : lean, one warp, waits+commits",           17, 2, 3, 3, 0, 3 | 256, 64},
      {"d64 : lean, one warp, no waits/commits",        17, 2, 3, 3, 0, 256, 64},
      {"d64 : lean, ksplit, one MMA of three",          17, 2, 3, 3, 0, 3 | 64 | 128 | 256 | 16, 64},
      {"d128: lean, ksplit, waits+commits",             4, 4, 3, 3, 0, 3 | 64 | 128 | 256, 128},
      {"d128: lean, ksplit, no waits/commits",          4, 4, 3, 3, 0, 64 | 128 | 256, 128},
      {"d128: lean, one warp, waits+commits",           4, 4, 3, 3, 0, 3 | 256, 128},
  };
  for (auto &e : cfgs) {
    Cfg c{};
    c.units = e.units; c.groups = e.groups; c.chunks = e.chunks; c.spc = e.spc; c.two = e.two; c.flags = e.flags; c.N = e.N;
    for (int i = 0; i < 32; ++i) { c.steps[i] = (uint32_t)(i * 260 + (i % 3)) | (130u << 16); c.steps64[i] = c.steps[i]; }
    c.desc_hi_opaque = 8u | (1u << 14);
    if (e.spc == 4) c.chunks = 4;
    bench<<<148, 224, 200 * 1024>>>(c, d);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    int n = c.units * c.groups * c.chunks * c.spc * ((e.flags & 16) ? 1 : 3) * (e.two ? 2 : 1);  // both warps' MMAs counted for dual issue
    printf("%-50s %7.1f cycles/MMA  (%d MMAs, %lld cycles, %s)\n", e.name, (double)mx / n, n, mx, cudaGetErrorString(err));
    if (err != cudaSuccess) return 1;
  }
  return 0;
}
