// mma_bench2.cu -- replicate conv_tc_kernel's issue pattern for a small-N layer (x-fold final conv: 3 groups per tile,
// 6 K steps x 3 MMAs per group, N = 32) in isolation, and bisect what costs time.  Timing only.
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
struct Cfg { int N, tiles, groups, steps, flags; };  // flags: 1 waits(complete barrier)+fence per group, 2 commit per group, 4 commit per tile to t_full
__global__ void __launch_bounds__(224, 1) bench(Cfg c, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done_bar, ready_bar, sink[8];
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u ^ ((uint32_t)i * 2654435761u & 0x03ff03ffu);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ready_bar)));
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sink[i])));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready_bar)) : "memory");  // phase 0 complete
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
  if (threadIdx.x < 32) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t desc_hi = 8u | (1u << 14);
    const uint32_t a_hi16 = smem_u32(smem) >> 4, a_lo16 = a_hi16 + (24576 >> 4);
    const uint32_t b0 = (smem_u32(smem + 100 * 1024) >> 4) | ((uint32_t)c.N << 16);
    const uint32_t b_step16 = 2u * c.N, b_lo16 = (uint32_t)c.steps * b_step16;
    long long t0 = clock64();
    uint32_t sink_i = 0;
    for (int tile = 0; tile < c.tiles; ++tile) {
      const uint32_t d0 = tmem + (tile & 1) * 256u;
      uint32_t acc = 0;
      for (int g = 0; g < c.groups; ++g) {
        if (c.flags & 1) {
          wait(&ready_bar, 0);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          wait(&ready_bar, 0);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (leader) {
#pragma unroll
          for (int st = 0; st < 8; ++st) {
            if (st < c.steps) {
              const uint32_t dl = (uint32_t)((st >> 1) * 4 + (st & 1) * 2) * 128u | (128u << 16);  // like (ri*CbG+2j)*pslab16 | lbo<<16
              const uint32_t bs = b0 + (uint32_t)st * b_step16;
              const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | (a_hi16 + dl), ad_lo = ((uint64_t)desc_hi << 32) | (a_lo16 + dl);
              const uint64_t bd_hi = ((uint64_t)desc_hi << 32) | bs, bd_lo = ((uint64_t)desc_hi << 32) | (bs + b_lo16);
              mma(d0, ad_hi, bd_hi, idesc, acc);
              mma(d0, ad_lo, bd_hi, idesc, 1);
              mma(d0, ad_hi, bd_lo, idesc, 1);
              acc = 1;
            }
          }
          if (c.flags & 2) { commit(&sink[sink_i & 7]); }
        }
        acc = 1;
        ++sink_i;
      }
      if ((c.flags & 4) && leader) commit(&sink[(sink_i++) & 7]);
    }
    if (leader) commit(&done_bar);
    wait(&done_bar, 0);
    long long t1 = clock64();
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long *d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char *name; Cfg c; } cfgs[] = {
      {"N32 3 groups x 6 steps, no sync",        {32, 50, 3, 6, 0}},
      {"N32 + commit per group",                 {32, 50, 3, 6, 2}},
      {"N32 + waits+fences per group",           {32, 50, 3, 6, 1}},
      {"N32 + waits+fences+commits (kernel)",    {32, 50, 3, 6, 7}},
      {"N32 1 group x 8 steps x many (kernel)",  {32, 200, 1, 8, 7}},
      {"N128 4 groups x 3 steps (res, kernel)",  {128, 50, 4, 3, 7}},
      {"N128 4 groups x 8 steps (kernel)",       {128, 50, 4, 8, 7}},
  };
  for (auto &e : cfgs) {
    bench<<<148, 224, 200 * 1024>>>(e.c, d);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    int n = e.c.tiles * e.c.groups * e.c.steps * 3;
    printf("%-42s %7.1f cycles/MMA  (%d MMAs, %s)\n", e.name, (double)mx / n, n, cudaGetErrorString(err));
    if (err != cudaSuccess) return 1;
  }
  return 0;
}
