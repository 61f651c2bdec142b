// occlusion.cuh -- the per-pixel forward/backward-flow consistency test of checkConsistency
// (consistencyChecker/consistencyChecker.cpp:99-125) in the reference's mixed float/double arithmetic, shared by the
// standalone kernel (consistency.cu) and the fused temporal-stage kernel (front.cu).  Bit-identical to the reference binary.
#pragma once
#include "fav_common.cuh"

namespace fav {

__device__ __forceinline__ float lerp_mixed(float alpha, float v0, float v1) {
  // (1.0 - alpha) * v0 + alpha * v1   with 1.0 a double literal: double*float -> double; alpha*v1 in float
  double a = __dmul_rn(__dsub_rn(1.0, (double)alpha), (double)v0);
  float b = __fmul_rn(alpha, v1);
  return (float)__dadd_rn(a, (double)b);
}

__device__ __forceinline__ uint8_t check_pixel(const float *__restrict__ f2u, const float *__restrict__ f2v,
                                               float u2, float v2, int ax, int ay, int W, int H,
                                               const float *__restrict__ structure, float structureAvg) {
  float bx = __fadd_rn((float)ax, u2);  // :101
  float by = __fadd_rn((float)ay, v2);
  int x1 = (int)floorf(bx), y1 = (int)floorf(by);  // :103-104
  int x2 = x1 + 1, y2 = y1 + 1;
  if (x1 < 0 || x2 >= W || y1 < 0 || y2 >= H) return 0;  // :107-108
  float alphaX = __fsub_rn(bx, (float)x1), alphaY = __fsub_rn(by, (float)y1);  // :109
  int64_t i11 = (int64_t)y1 * W + x1, i21 = i11 + 1, i12 = i11 + W, i22 = i12 + 1;
  float a = lerp_mixed(alphaX, __ldg(f2u + i11), __ldg(f2u + i21));  // :110
  float b = lerp_mixed(alphaX, __ldg(f2u + i12), __ldg(f2u + i22));
  float u = lerp_mixed(alphaY, a, b);
  a = lerp_mixed(alphaX, __ldg(f2v + i11), __ldg(f2v + i21));
  b = lerp_mixed(alphaX, __ldg(f2v + i12), __ldg(f2v + i22));
  float v = lerp_mixed(alphaY, a, b);
  float cx = __fadd_rn(bx, u), cy = __fadd_rn(by, v);  // :116-117
  float structureTerm = 0.f;
  if (structure) {  // :122-123
    float s = __fsub_rn(__fdiv_rn(structureAvg, 2.0f), __ldg(structure + (int64_t)ay * W + ax));
    structureTerm = __fmul_rn(__fdiv_rn(4.0f, structureAvg), fmaxf(0.0f, s));
  }
  float ex = __fsub_rn(cx, (float)ax), ey = __fsub_rn(cy, (float)ay);
  float lhs = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
  float mag = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(u2, u2), __fmul_rn(v2, v2)), __fmul_rn(u, u)),
                        __fmul_rn(v, v));
  double rhs = __dadd_rn(__dadd_rn(__dmul_rn(0.01, (double)mag), (double)structureTerm), (double)0.5f);  // :124
  return ((double)lhs >= rhs) ? 0 : 255;
}

}  // namespace fav
