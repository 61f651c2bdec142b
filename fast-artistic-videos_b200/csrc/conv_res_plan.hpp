// conv_res_plan.hpp -- host-only planning of the residual-block convolution kernel (conv_res.cu): eligibility, the
// cost-balanced tile table and the device job.  Pure functions (no CUDA runtime calls) shared by net.cu and the CPU
// emulator under tests/emu.
#pragma once
#include <algorithm>
#include <vector>

#include "conv_plan.hpp"
#include "conv_res.cuh"

namespace fav {

// relative cost of one tile of n 16-pixel granules (2 output rows): a tcgen05.mma with M = 128 takes N/2 = 8n cycles but
// never less than ~46 (issue / small-N floor, tools/mma_bench.cu); + a per-tile share of pipeline fill and epilogue hand-over
static inline int res_tile_cost(int n) { return std::max(8 * n, 46) + 7; }

// tiles of the flattened granule range [g0, g1) (row-pair major, Grow granules per row pair): cut at row-pair boundaries,
// pieces longer than 8 granules (128 pixels = the TMEM budget of a two-row unit) are split into near-equal parts
static inline int res_split_range(int g0, int g1, int Grow, std::vector<ResTile> *out) {
  int cost = 0;
  while (g0 < g1) {
    const int rp = g0 / Grow, row_end = std::min(g1, (rp + 1) * Grow), len = row_end - g0;
    const int k = (len + 7) / 8;
    for (int i = 0; i < k; ++i) {
      const int n = len / k + (i < len % k ? 1 : 0);
      if (out) out->push_back(ResTile{(int16_t)(2 * rp), (int16_t)((g0 % Grow) * 16), (int16_t)(n * 16), 0});
      cost += res_tile_cost(n);
      g0 += n;
    }
  }
  return cost;
}

struct ResPlan {
  std::vector<ResTile> tiles;
  std::vector<int> cta_first;  // [grid + 1]
  int grid = 0, max_cost = 0, sum_cost = 0;
};

// Equal shares of the Ho/2 x ceil(Wo/16) granules for `nctas` CTAs.  The ideal cut points k*G/n may move by up to D = 6 granules;
// a small dynamic program picks the offsets that minimise the most expensive CTA (it steers the cuts away from row-pair
// boundaries, where a cut a granule or two off would create a sliver tile at the 46-cycle floor).
static inline ResPlan plan_res_tiles(int Ho, int Wo, int nctas) {
  ResPlan pl;
  const int Grow = (Wo + 15) / 16, R = (Ho + 1) / 2, G = R * Grow;
  const int n = std::max(1, std::min(nctas, G));
  const int D = (G / n >= 6) ? 6 : 0, ND = 2 * D + 1;
  auto ideal = [&](int k) { return (int)(((long long)k * G + n / 2) / n); };
  auto bound = [&](int k, int j) {
    if (k == 0) return 0;
    if (k == n) return G;
    return std::min(G, std::max(0, ideal(k) + j - D));
  };
  // lexicographic objective (most expensive CTA, then total cost): a minimax alone leaves every CTA below the maximum free
  // to pick wasteful cuts
  const long long INF = 1ll << 60;
  auto key = [](int mx, long long sum) { return ((long long)mx << 32) + sum; };
  std::vector<long long> dp((size_t)(n + 1) * ND, INF);
  std::vector<int> from((size_t)(n + 1) * ND, -1);
  dp[D] = 0;  // boundary 0 is fixed
  for (int k = 1; k <= n; ++k)
    for (int j = 0; j < ND; ++j) {
      if ((k == n) && j != D) continue;
      const int e = bound(k, j);
      for (int i = 0; i < ND; ++i) {
        const long long prev = dp[(size_t)(k - 1) * ND + i];
        if (prev >= INF) continue;
        const int s = bound(k - 1, i);
        if (e <= s) continue;
        const int c = res_split_range(s, e, Grow, nullptr);
        const long long v = key(std::max((int)(prev >> 32), c), (prev & 0xffffffffll) + c);
        if (v < dp[(size_t)k * ND + j]) { dp[(size_t)k * ND + j] = v; from[(size_t)k * ND + j] = i; }
      }
    }
  std::vector<int> cut(n + 1);
  int j = D;
  for (int k = n; k >= 1; --k) {
    cut[k] = bound(k, j);
    j = from[(size_t)k * ND + j];
    if (j < 0) { j = D; }  // unreachable for G >= n; keeps the walk defined
  }
  cut[0] = 0;
  pl.grid = n;
  pl.cta_first.push_back(0);
  for (int k = 0; k < n; ++k) {
    const int c = res_split_range(cut[k], cut[k + 1], Grow, &pl.tiles);
    pl.max_cost = std::max(pl.max_cost, c);
    pl.sum_cost += c;
    pl.cta_first.push_back((int)pl.tiles.size());
  }
  return pl;
}

// the layers conv_res.cu covers: 3x3, stride 1, Cout = 128, Cin a multiple of 32, planned by build_phase_tables as the
// generic kind-0 phase with four channel blocks per stage (the table / weight image layout the kernel hard-codes)
static inline bool conv_res_eligible(const ConvDef &c, const ConvPhase &ph) {
  return !c.transposed && c.k == 3 && c.stride == 1 && c.in_stride == 1 && c.cout == 128 && c.Cb % kResCbG == 0 &&
         ph.kind == 0 && ph.CbG == kResCbG && ph.nrg == 1 && ph.rows_per_group == 3 && ph.pslab16 == kResPslab &&
         (int)ph.steps.size() == kResSteps && ph.spc == kResSpc && ph.Npad == 128 && !ph.rf_R && !ph.pf &&
         ph.rows.size() == 3 && ph.rows[1] == ph.rows[0] + 1 && ph.rows[2] == ph.rows[0] + 2 && ph.dxmax - ph.dxmin == 2;
}

// device job of one layer (tiles / cta_first / b / bias / raw / stats pointers filled by the caller)
static inline int fill_res_job(const ConvDef &c, const ConvPhase &ph, const Operand &in, ResJob &j) {
  int Ho, Wo;
  conv_out_size(c, in.H, in.W, &Ho, &Wo);
  memset(&j, 0, sizeof(j));
  j.a_hi = reinterpret_cast<const uint4 *>(in.hi); j.a_lo = reinterpret_cast<const uint4 *>(in.lo);
  j.a_Cb = in.Cb; j.a_slab16 = in.slab16();
  j.in_row0 = in.padT + ph.rows[0]; j.in_col0 = in.padL + ph.dxmin;
  j.Ho = Ho; j.Wo = Wo; j.ngroups = c.Cb / kResCbG;
  for (int i = 0; i < kResSteps; ++i) j.steps[i] = (uint32_t)ph.steps[i].a_off16 | ((uint32_t)ph.steps[i].lbo16 << 16);
  // every bulk copy must stay inside the operand allocation: last row pair, last granule, halo
  const int R = (Ho + 1) / 2, Grow = (Wo + 15) / 16;
  const int64_t max_row = (int64_t)(2 * (R - 1) + 3) + j.in_row0;
  const int64_t last16 = (max_row * in.Cb + in.Cb - 1) * (int64_t)in.slab16() + j.in_col0 + (int64_t)Grow * 16 + 2;
  if (in.parity || j.in_row0 < 0 || j.in_col0 < 0 || max_row >= in.Hs || last16 > (int64_t)in.elems16) {
    set_error("conv %s: operand bounds (residual kernel)", c.name.c_str());
    return FAV_ERR_INVALID;
  }
  return FAV_OK;
}

}  // namespace fav
