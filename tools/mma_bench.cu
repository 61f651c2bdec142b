// mma_bench.cu -- microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, cta_group::1) for different shared-memory
// operand layouts (no-swizzle/interleaved vs 128B swizzle), N, and descriptor strides.  Timing only (operands are
// zeros).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct Cfg { int N; int layout; uint32_t a_lbo16, a_sbo16, b_lbo16, b_sbo16; int n_mma; int same_acc; int a_stride_bytes; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  if (threadIdx.x < 32) {  // whole warp runs the loop (uniform control flow); one elected lane issues
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 100 * 1024);
    auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo & 0x3FFF) << 16) | ((uint64_t)(sbo & 0x3FFF) << 32) |
             (1ull << 46) | ((uint64_t)c.layout << 61);
    };
    long long t0 = clock64();
    const uint64_t ad0 = desc(a0, c.a_lbo16, c.a_sbo16), bd0 = desc(b0, c.b_lbo16, c.b_sbo16);
    const uint32_t astep = (uint32_t)c.a_stride_bytes >> 4;
#pragma unroll 8
    for (int i = 0; i < c.n_mma; ++i) {
      uint64_t ad = ad0 + (uint64_t)((i & 7) * astep), bd = bd0 + (uint64_t)((i & 3) * 2);
      if (leader) mma(tmem + (c.same_acc ? 0 : (i & 1) * 256), ad, bd, idesc, i > 0);
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
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
      // name                       N  layout  a_lbo a_sbo b_lbo b_sbo  n   same stride
      {"noswz N128 mine(lbo2080)", {128, 0, 130, 8, 128, 8, 2000, 1, 16}},
      {"noswz N128 dense(lbo128)", {128, 0, 8, 16, 8, 16, 2000, 1, 16}},     // core matrices of a K pair adjacent
      {"noswz N32  mine",          {32, 0, 130, 8, 32, 8, 2000, 1, 16}},
      {"noswz N16  mine",          {16, 0, 130, 8, 16, 8, 2000, 1, 16}},
      {"noswz N128 lbo=1(16B)",    {128, 0, 1, 8, 128, 8, 2000, 1, 16}},
      {"sw128 N128",               {128, 2, 1, 64, 1, 64, 2000, 1, 32}},
      {"sw128 N32",                {32, 2, 1, 64, 1, 64, 2000, 1, 32}},
      {"sw128 N256",               {256, 2, 1, 64, 1, 64, 2000, 1, 32}},
      {"sw64  N128",               {128, 4, 1, 32, 1, 32, 2000, 1, 32}},
      {"sw32  N128",               {128, 6, 1, 16, 1, 16, 2000, 1, 32}},
      {"noswz N128 mine alt-acc",  {128, 0, 130, 8, 128, 8, 2000, 0, 16}},
      {"noswz N256 mine",          {256, 0, 130, 8, 256, 8, 2000, 1, 16}},
      {"noswz N64 mine",           {64, 0, 130, 8, 64, 8, 2000, 1, 16}},
  };
  for (int g = 1; g <= 148; g += 147) {
    for (auto &e : cfgs) {
      bench<<<g, 128, 200 * 1024>>>(e.c, d);
      cudaError_t err = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, d, g * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < g; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("grid %3d  %-28s  %7.1f cycles/MMA  (%s)\n", g, e.name, (double)mx / e.c.n_mma, cudaGetErrorString(err));
      if (err != cudaSuccess) return 1;
    }
  }
  return 0;
}
