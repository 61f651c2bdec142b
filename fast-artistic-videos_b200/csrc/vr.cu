// vr.cu -- VR (cube-map) post-processing kernels, SURVEY.md §8 a-V / f-4:
//   median filter                 utils.median_filter            fast_artistic_video/utils.lua:151-159
//       (r x r windows by unfold, median over the r*r values = lower median, VALID region (H-r+1) x (W-r+1);
//        the reference round-trips to the CPU for it because CudaTensor has no median)
//   fused 4-way border blend      combineSides + blend           fast_artistic_video_vr.lua:146-152, 454-509
//       out = base * (1 - mask) + mask * sum_{i<4} warp(rot_i(side_i), map_i) / div
//       -- the reference runs 4 warp launches, 4 rotations (index copies), 4 divisions, 3 adds, 2 cmuls, 1 add per face;
//       here the rotation is an index mapping inside the gather and everything is one kernel.
// Arithmetic follows the reference order with round-to-nearest intrinsics (bit-exact vs the oracle composition).
#include "fav_common.cuh"

namespace fav {

// ---- median ---------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) median_kernel(const float *__restrict__ in, float *__restrict__ out, int C, int H, int W) {
  const int Wo = W - R + 1, Ho = H - R + 1;
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, c = blockIdx.z;
  if (x >= Wo || y >= Ho) return;
  float v[R * R];
  const float *p = in + ((int64_t)c * H + y) * W + x;
#pragma unroll
  for (int dy = 0; dy < R; ++dy)
#pragma unroll
    for (int dx = 0; dx < R; ++dx) v[dy * R + dx] = __ldg(p + (int64_t)dy * W + dx);
  // partial selection sort up to the (lower) median index: torch.median returns element (n-1)/2 of the sorted values
  constexpr int N = R * R, K = (N - 1) / 2;
#pragma unroll
  for (int i = 0; i <= K; ++i) {
#pragma unroll
    for (int j = i + 1; j < N; ++j) {
      float a = v[i], b = v[j];
      v[i] = fminf(a, b);
      v[j] = fmaxf(a, b);
    }
  }
  out[((int64_t)c * Ho + y) * Wo + x] = v[K];
}

// ---- fused border blend -------------------------------------------------------------------------------------------
struct BlendSide {
  const float *img;   // [3,S,S] stylized neighbour face
  const float *map;   // [2,S,S] perspective warp map (dy,dx), sentinel 99999 outside the border strip
  int rot;            // 0 none, 1 rotate90, 2 rotateMinus90, 3 rotate180   (fast_artistic_video_vr.lua:134-144)
};
struct BlendArgs {
  BlendSide s[4];
  const float *base;      // [3,S,S]
  const float *div;       // [S,S] mask_all_div
  const float *mask;      // [S,S] blend mask (grad_mask_all)
  const float *anti;      // [S,S] 1 - grad_mask_all as the reference casts it from double (:456), or null
  float *out;             // [3,S,S]
  int S;
};

// value of rot(img)[c][y][x] for a square S x S face
__device__ __forceinline__ float rot_fetch(const float *__restrict__ img, int rot, int c, int y, int x, int S) {
  int sy, sx;
  if (rot == 0) { sy = y; sx = x; }
  else if (rot == 1) { sy = x; sx = S - 1 - y; }       // rotate90:      R[y][x] = t[x][S-1-y]
  else if (rot == 2) { sy = S - 1 - x; sx = y; }       // rotateMinus90: R[y][x] = t[S-1-x][y]
  else { sy = S - 1 - y; sx = S - 1 - x; }             // rotate180
  return __ldg(img + ((int64_t)c * S + sy) * S + sx);
}

__global__ void __launch_bounds__(256) vr_blend_kernel(const __grid_constant__ BlendArgs a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int S = a.S;
  if (x >= S || y >= S) return;
  const int64_t o = (int64_t)y * S + x, SS = (int64_t)S * S;
  const float m = __ldg(a.mask + o), dv = __ldg(a.div + o);
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float dy = __ldg(a.s[i].map + o), dx = __ldg(a.s[i].map + SS + o);
    // bilinear sample exactly as BilinearSamplerBDHW.cu:72-108 (per-corner zero fill)
    float yf = __fadd_rn(dy, (float)y), xf = __fadd_rn(dx, (float)x);
    float fy = floorf(yf), fx = floorf(xf);
    int y0 = (int)fy, x0 = (int)fx;
    float wx = __fsub_rn(1.0f, __fsub_rn(xf, (float)x0)), wy = __fsub_rn(1.0f, __fsub_rn(yf, (float)y0));
    bool xin0 = x0 >= 0 && x0 <= S - 1, xin1 = x0 + 1 >= 0 && x0 + 1 <= S - 1;
    bool yin0 = y0 >= 0 && y0 <= S - 1, yin1 = y0 + 1 >= 0 && y0 + 1 <= S - 1;
    float omx = __fsub_rn(1.0f, wx), omy = __fsub_rn(1.0f, wy);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float vtl = (xin0 && yin0) ? rot_fetch(a.s[i].img, a.s[i].rot, c, y0, x0, S) : 0.f;
      float vtr = (xin1 && yin0) ? rot_fetch(a.s[i].img, a.s[i].rot, c, y0, x0 + 1, S) : 0.f;
      float vbl = (xin0 && yin1) ? rot_fetch(a.s[i].img, a.s[i].rot, c, y0 + 1, x0, S) : 0.f;
      float vbr = (xin1 && yin1) ? rot_fetch(a.s[i].img, a.s[i].rot, c, y0 + 1, x0 + 1, S) : 0.f;
      float v = bilinear_ref_blend(wx, wy, omx, omy, vtl, vtr, vbl, vbr);
      float q = __fdiv_rn(v, dv);                      // torch.cdiv(side, divisor)   (combineSides :147-151)
      acc[c] = i == 0 ? q : __fadd_rn(acc[c], q);
    }
  }
  const float am = a.anti ? __ldg(a.anti + o) : __fsub_rn(1.0f, m);  // anti_mask = 1 - grad_mask_all  (:456)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float b = __ldg(a.base + c * SS + o);
    a.out[c * SS + o] = __fadd_rn(__fmul_rn(b, am), __fmul_rn(acc[c], m));  // :466
  }
}

}  // namespace fav

using namespace fav;

extern "C" {

int fav_median_filter(const float *in, float *out, int C, int H, int W, int r, void *stream) {
  FAV_REQUIRE(in && out, "median_filter: null tensor");
  FAV_REQUIRE(r == 1 || r == 3 || r == 5, "median_filter: r must be 1, 3 or 5");
  FAV_REQUIRE(C > 0 && H >= r && W >= r, "median_filter: image smaller than the window");
  FAV_TRY(require_device());
  cudaStream_t st = (cudaStream_t)stream;
  dim3 block(32, 8), grid(ceil_div(W - r + 1, 32), ceil_div(H - r + 1, 8), C);
  if (r == 1) median_kernel<1><<<grid, block, 0, st>>>(in, out, C, H, W);
  else if (r == 3) median_kernel<3><<<grid, block, 0, st>>>(in, out, C, H, W);
  else median_kernel<5><<<grid, block, 0, st>>>(in, out, C, H, W);
  return post_launch("median_filter");
}

int fav_vr_blend_sides(const float *base, const float *const sides[4], const float *const maps[4], const int rot[4],
                       const float *div, const float *mask, const float *anti_mask, float *out, int S, void *stream) {
  FAV_REQUIRE(base && sides && maps && rot && div && mask && out, "vr_blend_sides: null argument");
  FAV_REQUIRE(S > 1, "vr_blend_sides: empty face");
  BlendArgs a;
  for (int i = 0; i < 4; ++i) {
    FAV_REQUIRE(sides[i] && maps[i] && rot[i] >= 0 && rot[i] <= 3, "vr_blend_sides: bad side %d", i);
    a.s[i].img = sides[i]; a.s[i].map = maps[i]; a.s[i].rot = rot[i];
  }
  a.base = base; a.div = div; a.mask = mask; a.anti = anti_mask; a.out = out; a.S = S;
  FAV_TRY(require_device());
  dim3 block(32, 8), grid(ceil_div(S, 32), ceil_div(S, 8));
  vr_blend_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a);
  return post_launch("vr_blend_sides");
}
}
