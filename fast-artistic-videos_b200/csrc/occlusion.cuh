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

// The test in two halves so that a caller can have the gathers of several pixels in flight before the (long, dependent)
// mixed-precision arithmetic of the first one starts.
struct CheckTaps {
  float bx, by, alphaX, alphaY;
  float u11, u21, u12, u22, v11, v21, v12, v22;
  bool inside;
};

__device__ __forceinline__ CheckTaps check_load(const float *__restrict__ f2u, const float *__restrict__ f2v, float u2, float v2,
                                                int ax, int ay, int W, int H) {
  CheckTaps t;
  t.bx = __fadd_rn((float)ax, u2);  // :101
  t.by = __fadd_rn((float)ay, v2);
  const int x1 = (int)floorf(t.bx), y1 = (int)floorf(t.by);  // :103-104
  const int x2 = x1 + 1, y2 = y1 + 1;
  t.inside = !(x1 < 0 || x2 >= W || y1 < 0 || y2 >= H);  // :107-108
  t.alphaX = __fsub_rn(t.bx, (float)x1); t.alphaY = __fsub_rn(t.by, (float)y1);  // :109
  t.u11 = t.u21 = t.u12 = t.u22 = t.v11 = t.v21 = t.v12 = t.v22 = 0.f;
  if (t.inside) {
    const int64_t i11 = (int64_t)y1 * W + x1, i21 = i11 + 1, i12 = i11 + W, i22 = i12 + 1;
    t.u11 = __ldg(f2u + i11); t.u21 = __ldg(f2u + i21); t.u12 = __ldg(f2u + i12); t.u22 = __ldg(f2u + i22);
    t.v11 = __ldg(f2v + i11); t.v21 = __ldg(f2v + i21); t.v12 = __ldg(f2v + i12); t.v22 = __ldg(f2v + i22);
  }
  return t;
}

__device__ __forceinline__ uint8_t check_eval(const CheckTaps &t, float u2, float v2, int ax, int ay, int W,
                                              const float *__restrict__ structure, float structureAvg) {
  if (!t.inside) return 0;
  float a = lerp_mixed(t.alphaX, t.u11, t.u21);  // :110
  float b = lerp_mixed(t.alphaX, t.u12, t.u22);
  const float u = lerp_mixed(t.alphaY, a, b);
  a = lerp_mixed(t.alphaX, t.v11, t.v21);
  b = lerp_mixed(t.alphaX, t.v12, t.v22);
  const float v = lerp_mixed(t.alphaY, a, b);
  const float cx = __fadd_rn(t.bx, u), cy = __fadd_rn(t.by, v);  // :116-117
  float structureTerm = 0.f;
  if (structure) {  // :122-123
    float s = __fsub_rn(__fdiv_rn(structureAvg, 2.0f), __ldg(structure + (int64_t)ay * W + ax));
    structureTerm = __fmul_rn(__fdiv_rn(4.0f, structureAvg), fmaxf(0.0f, s));
  }
  const float ex = __fsub_rn(cx, (float)ax), ey = __fsub_rn(cy, (float)ay);
  const float lhs = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
  const float mag = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(u2, u2), __fmul_rn(v2, v2)), __fmul_rn(u, u)),
                              __fmul_rn(v, v));
  const double rhs = __dadd_rn(__dadd_rn(__dmul_rn(0.01, (double)mag), (double)structureTerm), (double)0.5f);  // :124
  return ((double)lhs >= rhs) ? 0 : 255;
}

__device__ __forceinline__ uint8_t check_pixel(const float *__restrict__ f2u, const float *__restrict__ f2v,
                                               float u2, float v2, int ax, int ay, int W, int H,
                                               const float *__restrict__ structure, float structureAvg) {
  const CheckTaps t = check_load(f2u, f2v, u2, v2, ax, ay, W, H);
  return check_eval(t, u2, v2, ax, ay, W, structure, structureAvg);
}

}  // namespace fav
