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

struct Cfg { int N; int layout; uint32_t a_lbo16, a_sbo16, b_lbo16, b_sbo16; int n_mma; int same_acc; int a_stride_bytes; int every; int what; int shift; int b_stride_bytes; int M; };

__global__ void __launch_bounds__(224, 1) bench(Cfg c, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = (c.what & 8) ? (0x3c003c00u ^ ((uint32_t)i * 2654435761u & 0x03ff03ffu)) : 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)((c.M ? c.M : 128) >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 100 * 1024);
    auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo & 0x3FFF) << 16) | ((uint64_t)(sbo & 0x3FFF) << 32) |
             (1ull << 46) | ((uint64_t)c.layout << 61);
    };
    long long t0 = clock64();
    const uint64_t ad0 = desc(a0, c.a_lbo16, c.a_sbo16), bd0 = desc(b0, c.b_lbo16, c.b_sbo16);
    const uint32_t astep = (uint32_t)c.a_stride_bytes >> 4, bstep = (uint32_t)c.b_stride_bytes >> 4;
#pragma unroll 8
    for (int i = 0; i < c.n_mma; ++i) {
      uint64_t ad = ad0 + (uint64_t)((i & 15) * astep), bd = bd0 + (uint64_t)((i & 7) * bstep);
      if (leader) mma(tmem + (c.same_acc ? 0 : (i & 1) * 256), ad, bd, idesc, i > 0);
      if (c.every && (i & (c.every - 1)) == c.every - 1) {
        if ((c.what & 1) && leader)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[(i >> c.shift) & 7])) : "memory");
        if (c.what & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (c.what & 4) {  // wait for the commit issued 4 intervals ago (ring of 8 barriers, like a weight-slot ring)
          int k = i >> c.shift;
          if (k >= 4) {
            uint32_t done = 0; uint32_t par = ((k - 4) >> 3) & 1;
            while (!done)
              asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bars[(k - 4) & 7])), "r"(par) : "memory");
          }
        }
      }
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  else if (c.what & 16) {
    // polling warps: wait for the final commit exactly like the conv kernel's epilogue / producer warps do
    uint32_t done = 0;
    const bool one_lane = (c.what & 32) != 0;
    if (!one_lane || (threadIdx.x & 31) == 0) {
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        if ((c.what & 64) && !done) __nanosleep(200);
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long *d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char *name; Cfg c; } cfgs[] = {
      {"N32  no pollers",                      {32, 0, 130, 8, 32, 8, 4096, 1, 4096, 0, 8, 0, 1024}},
      {"N32  6 warps polling (32 lanes)",      {32, 0, 130, 8, 32, 8, 4096, 1, 4096, 0, 8 | 16, 0, 1024}},
      {"N32  6 warps polling (1 lane)",        {32, 0, 130, 8, 32, 8, 4096, 1, 4096, 0, 8 | 16 | 32, 0, 1024}},
      {"N32  6 warps polling + nanosleep",     {32, 0, 130, 8, 32, 8, 4096, 1, 4096, 0, 8 | 16 | 64, 0, 1024}},
      {"N128 no pollers",                      {128, 0, 130, 8, 128, 8, 4096, 1, 4096, 0, 8, 0, 4096}},
      {"N128 6 warps polling (32 lanes)",      {128, 0, 130, 8, 128, 8, 4096, 1, 4096, 0, 8 | 16, 0, 4096}},
      {"N128 6 warps polling + nanosleep",     {128, 0, 130, 8, 128, 8, 4096, 1, 4096, 0, 8 | 16 | 64, 0, 4096}},
      {"N128 A start +16B steps",              {128, 0, 130, 8, 128, 8, 4096, 1, 16, 0, 8, 0, 4096}},
      {"N128 A start +32B steps",              {128, 0, 130, 8, 128, 8, 4096, 1, 32, 0, 8, 0, 4096}},
      {"N128 A start +64B steps",              {128, 0, 130, 8, 128, 8, 4096, 1, 64, 0, 8, 0, 4096}},
      {"N128 A start +128B steps",             {128, 0, 130, 8, 128, 8, 4096, 1, 128, 0, 8, 0, 4096}},
      {"N128 A start +2080B steps",            {128, 0, 130, 8, 128, 8, 4096, 1, 2080, 0, 8, 0, 4096}},
      {"N128 A lbo 128 (aligned) +4096",       {128, 0, 128, 8, 128, 8, 4096, 1, 4096, 0, 8, 0, 4096}},
      {"N128 A lbo 128, +16B steps",           {128, 0, 128, 8, 128, 8, 4096, 1, 16, 0, 8, 0, 4096}},
      {"N128 two accumulators, +16B",          {128, 0, 130, 8, 128, 8, 4096, 0, 16, 0, 8, 0, 4096}},
      {"N128 B start +16B steps",              {128, 0, 130, 8, 128, 8, 4096, 1, 4096, 0, 8, 0, 16}},
      {"N256 A +4096",                         {256, 0, 130, 8, 256, 8, 4096, 1, 4096, 0, 8, 0, 0}},
      {"N256 A +16B steps",                    {256, 0, 130, 8, 256, 8, 4096, 1, 16, 0, 8, 0, 0}},
      {"N160 A +16B steps",                    {160, 0, 130, 8, 160, 8, 4096, 1, 16, 0, 8, 0, 0}},
      {"N160 aligned",                         {160, 0, 128, 8, 160, 8, 4096, 1, 4096, 0, 8, 0, 0}},
      {"M64 N128",                             {128, 0, 130, 8, 128, 8, 4096, 1, 4096, 0, 8, 0, 4096, 64}},
      {"M64 N256",                             {256, 0, 130, 8, 256, 8, 4096, 1, 4096, 0, 8, 0, 0, 64}},
      {"M64 N64",                              {64, 0, 130, 8, 64, 8, 4096, 1, 4096, 0, 8, 0, 0, 64}},
      {"M64 N128 two accumulators",            {128, 0, 130, 8, 128, 8, 4096, 0, 4096, 0, 8, 0, 4096, 64}},
      // the narrow layers of conv_tc (r02): N = 64 / 128 with the patch strides of the stride-2 and transposed plans
      {"N64  aligned",                         {64, 0, 128, 8, 64, 8, 4096, 1, 4096, 0, 8, 0, 2048}},
      {"N64  A lbo130 +16B steps",             {64, 0, 130, 8, 64, 8, 4096, 1, 16, 0, 8, 0, 2048}},
      {"N64  A lbo257 +16B steps (d64)",       {64, 0, 257, 8, 64, 8, 4096, 1, 16, 0, 8, 0, 2048}},
      {"N64  A lbo257 +16B, two accumulators", {64, 0, 257, 8, 64, 8, 4096, 0, 16, 0, 8, 0, 2048}},
      {"N64  A lbo257 commit/8 + ring wait",   {64, 0, 257, 8, 64, 8, 4096, 1, 16, 8, 8 | 1 | 4, 3, 2048}},
      {"N64  A lbo257 commit/8, pollers",      {64, 0, 257, 8, 64, 8, 4096, 1, 16, 8, 8 | 1 | 4 | 16, 3, 2048}},
      {"N128 A lbo257 +16B steps (d128)",      {128, 0, 257, 8, 128, 8, 4096, 1, 16, 0, 8, 0, 4096}},
      {"N128 A lbo129 +16B steps (u32)",       {128, 0, 129, 8, 128, 8, 4096, 1, 16, 0, 8, 0, 4096}},
      {"N128 A lbo129 commit/8 + ring wait",   {128, 0, 129, 8, 128, 8, 4096, 1, 16, 8, 8 | 1 | 4, 3, 4096}},
  };
  for (int g = 148; g <= 148; g += 147) {
    for (auto &e : cfgs) {
      bench<<<g, 224, 200 * 1024>>>(e.c, d);
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
