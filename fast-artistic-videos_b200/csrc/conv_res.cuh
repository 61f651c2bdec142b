// conv_res.cuh -- the residual-block convolution (3x3, stride 1, Cout = 128: 67 % of the net's FLOPs,
// fast_artistic_video/models_video.lua:20,32 inside build_res_block :41-53) as its own tcgen05 kernel with the GEMM roles
// SWAPPED relative to conv_tc.cu:
//     D[128 couts (TMEM lanes), nt pixels (TMEM columns)] += W[128 couts, K] * P[nt pixels, K]^T
// The weights are the M = 128 operand, the pixels the N operand.  tcgen05.mma costs max(M,128) * N / 256 cycles, so with the
// pixels along N a tile can be ANY multiple of 16 pixels wide at proportional cost (M is fixed at 128: a half-empty 128-pixel
// M tile costs as much as a full one -- the 14-20 % geometry loss of the first design).  Consequences:
//   * work = (row pair, 16-pixel granule range) tiles from a host-built table; every CTA gets an equal share of granules
//     (cost-balanced cuts, plan_res_tiles) instead of "2 or 3 fixed 128-pixel units";
//   * an epilogue thread owns ONE output channel (TMEM lane) and 16 consecutive pixels per tcgen05.ld: the InstanceNorm
//     statistics are plain per-thread sums (no shuffle butterflies), the raw output is PLANAR fp32 [C][Hp][Wp] written with
//     256-bit stores;
//   * the shared-memory images of patch and weights are the same as in conv_tc.cu (canonical no-swizzle K-major core
//     matrices), only the two matrix descriptors trade places.
#pragma once
#include "conv.cuh"

namespace fav {

constexpr int kResMaxNt = 128;      // widest tile: 2 rows x 128 columns x 2 TMEM stages = 512 columns
constexpr int kResCbG = 2;          // channel blocks per patch stage (16 input channels = one K16 step per tap): four 33 KB
                                    // stages in flight hide the fill latency of the norm-on-load producers (one slab per
                                    // producer warp and stage) and let the first MMA start after 33 KB instead of 66 KB
constexpr int kResPslab = kResMaxNt + 2;  // patch slab pitch in pixels (3x3: one halo pixel each side)
constexpr int kResSteps = 9;        // K16 steps per channel group: 9 taps x 1 channel-block pair
constexpr int kResSpc = 3;          // K16 steps per weight chunk (24 KB)
constexpr int kResChunks = kResSteps / kResSpc;

struct ResTile {
  int16_t y;    // first output row of the row pair
  int16_t x0;   // first output pixel (multiple of 16)
  int16_t nt;   // pixels (multiple of 16, <= kResMaxNt)
  int16_t pad_;
};

struct ResJob {
  // input: fp16 hi/lo operand (nl == 0) -- storage row of tap row ky for output row y: y + ky + in_row0; pixel: x + kx + in_col0
  const uint4 *a_hi, *a_lo;
  int a_Cb, a_slab16, in_row0, in_col0;
  // input: PLANAR raw fp32 of the previous convolution, normalised on load (nl == 1; InstanceNorm + ReLU + hi/lo split in the
  // producer warps, same arithmetic as in_apply_kernel).  Logical input pixel of tap (ky,kx): (y + ky - nl_pad, x + kx - nl_pad)
  int nl, nl_pad, nl_Hp, nl_Wp, nl_H, nl_W, nl_relu, nl_C;
  const float *nl_raw;
  const double *nl_sums;
  const float *nl_gamma, *nl_beta;
  double nl_inv_count, nl_eps;
  int Ho, Wo, ngroups;        // ngroups = Cin / 16
  uint32_t steps[kResSteps];  // KStep words (a_off16 | lbo16 << 16) of one channel group, order = weight packing order
  const uint4 *b;             // packed weights [group][chunk][hi|lo][step][k-half][128 couts][8 ch] (conv_plan.hpp, Npad = 128)
  const float *bias;
  float *raw;                 // planar fp32 output [128][raw_Hp][raw_Wp]
  int raw_Hp, raw_Wp;
  double *stats;              // [2][128] sum / sum of squares (double atomics), or null
  const ResTile *tiles;       // device
  const int *cta_first;       // device, [grid + 1]
  int grid;
  unsigned long long *trace;  // diagnostics (fav_debug_set_trace), layout as in conv_tc.cu
};

int launch_conv_res(const ResJob &job, cudaStream_t st);
size_t conv_res_smem_bytes(int b_slots);
int conv_res_slots();

}  // namespace fav
