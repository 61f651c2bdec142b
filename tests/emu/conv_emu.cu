// conv_emu.cu -- TEST INFRASTRUCTURE: a CPU emulation of conv_tc_kernel's ADDRESSING (bulk-copy segments, patch
// stages, K-step tables, canonical no-swizzle K-major UMMA operand layout, packed weight chunks, output
// placement) driven by the very same host planning code the product uses (csrc/conv_plan.hpp).  It checks the
// tables against a direct convolution, so that table / packing bugs are caught here on the CPU and GPU time is
// spent on hardware semantics only.  Not part of libfav_b200.so.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <random>
#include <vector>

#include "../../fast-artistic-videos_b200/csrc/conv_plan.hpp"
#include "../../fast-artistic-videos_b200/csrc/conv_res_plan.hpp"

namespace fav {
static char g_err[512];
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
std::atomic<uint64_t> g_launches{0};
// mirror of conv_tc.cu: kNA = 2 patch stages, up to 16 weight slots (resident when every chunk fits)
static size_t tc_fixed_smem(const ConvJob &job) {
  return (size_t)job.a_stages * 2 * job.stage16 * 16 + 640 + (job.xfold_kw ? (size_t)(job.mt >= 2 ? 2 : 1) * 128 * 33 * 4 : (size_t)(256 + 2048) * 4);
}
size_t conv_tc_smem_bytes(const ConvJob &job) { return tc_fixed_smem(job) + (size_t)job.b_slots * job.chunk16 * 16; }
void conv_tc_choose_slots(ConvJob &job) {
  const size_t budget = 224 * 1024, chunk = (size_t)job.chunk16 * 16;
  const int total = job.rf_R ? job.rf_steps : (job.pf ? job.ngroups * 4 : job.ngroups * job.nchunks);
  job.a_stages = 2;
  {
    ConvJob t = job;
    for (int n = 4; n > 2; --n) {
      t.a_stages = n;
      if (total <= 16 && tc_fixed_smem(t) + total * chunk <= budget) { job.a_stages = n; break; }
    }
  }
  const size_t fixed = tc_fixed_smem(job);
  if (total <= 16 && fixed + total * chunk <= budget) { job.b_resident = 1; job.b_slots = total; }
  else { job.b_resident = 0; int n = (int)((budget - fixed) / chunk); job.b_slots = n > 16 ? 16 : (n < 2 ? 2 : n); }
}
}  // namespace fav

using namespace fav;

extern "C" const char *emu_last_error() { return fav::g_err; }

// returns 0 on success; max_err = max |emulated - direct|, max_ref = max |direct|
extern "C" int emu_conv_check(int cin, int cout, int k, int stride, int pad, int transposed, int adj, int H, int W,
                              unsigned seed, double *max_err, double *max_ref, int *smem_bytes, int *n_mma) {
  ConvDef c;
  init_conv_def(c, "emu", cin, cout, k, stride, pad, transposed != 0, adj);
  build_phases(c);
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  std::vector<float> w((size_t)cin * cout * k * k), x((size_t)cin * H * W);
  for (auto &v : w) v = U(rng) * 0.1f;
  for (auto &v : x) v = U(rng) * 3.f;
  // operand (host mirror)
  Operand op = operand_geometry(cin, H, W, &c);
  std::vector<uint16_t> hi(op.elems16 * 8, 0), lo(op.elems16 * 8, 0);
  std::vector<float> xq((size_t)cin * H * W);
  for (int ci = 0; ci < cin; ++ci)
    for (int y = 0; y < H; ++y)
      for (int xx = 0; xx < W; ++xx) {
        float v = x[((size_t)ci * H + y) * W + xx];
        uint16_t hb = f2h_bits(v), lb = f2h_bits(v - h2f_bits(hb));
        int64_t o = op.off16(op.padT + y, ci / 8, op.padL + xx) * 8 + ci % 8;
        hi[o] = hb; lo[o] = lb;
        xq[((size_t)ci * H + y) * W + xx] = h2f_bits(hb) + h2f_bits(lb);
      }
  int Ho, Wo;
  conv_out_size(c, H, W, &Ho, &Wo);
  std::vector<double> out((size_t)cout * Ho * Wo, 0.0), ref((size_t)cout * Ho * Wo, 0.0);
  std::vector<char> written((size_t)Ho * Wo, 0);
  int mma_count = 0;
  size_t smem_max = 0;
  std::vector<ConvPhase *> todo;
  if (c.has_fold) todo.push_back(&c.fold);
  else for (ConvPhase &p : c.phases) todo.push_back(&p);
  for (ConvPhase *php : todo) {
    ConvPhase &ph = *php;
    if (!ph.pf && build_phase_tables(c, ph) != FAV_OK) return 1;
    std::vector<uint16_t> pk = ph.pf ? pack_phase_fold(c, ph, w) : pack_phase_weights(c, ph, w);
    ConvJob j;
    if (fill_conv_job(c, ph, op, j) != FAV_OK) return 2;
    conv_tc_choose_slots(j);
    if (j.b_slots < 2 && !j.b_resident) { set_error("weight ring too small"); return 9; }
    smem_max = std::max(smem_max, conv_tc_smem_bytes(j));
    if (conv_tc_smem_bytes(j) > 227 * 1024) { set_error("smem budget"); return 3; }
    const int Npad = j.Npad;
    std::vector<uint16_t> st_hi((size_t)j.stage16 * 8), st_lo((size_t)j.stage16 * 8);
    std::vector<double> acc(j.rf_R ? (size_t)kTileM * 512 : (size_t)2 * kTileM * Npad);
    std::vector<double> acc2(acc.size());  // K-split: the second issuing warp's accumulator (TMEM columns +128)
    if (j.ksplit && (j.rf_R ? j.rf_R * j.rf_nblk : Npad) > 128) { set_error("ksplit needs <= 128 columns"); return 10; }
    if (j.ksplit && j.mt != 1 && !j.rf_R) { set_error("ksplit with mt = 2"); return 10; }
    for (int tile = 0; tile < j.ntiles; ++tile) {
      const int y = (tile / j.tiles_x) * j.mt, x0 = (tile % j.tiles_x) * j.tile_dx;
      const double poison = (j.rf_R || j.ksplit) ? 1e30 : 0.0;  // stale TMEM must be overwritten, not accumulated
      std::fill(acc.begin(), acc.end(), poison);
      std::fill(acc2.begin(), acc2.end(), poison);
      bool first_w[2] = {true, true};  // K-split: first MMA of each issuing warp in the unit (accumulate = 0)
      int sc = 0;                       // K-step counter of the unit
      for (int g = 0; g < j.ngroups; ++g) {
        // A producer
        std::fill(st_hi.begin(), st_hi.end(), (uint16_t)0x7e00);  // NaN poison: reading an unloaded byte is a bug
        std::fill(st_lo.begin(), st_lo.end(), (uint16_t)0x7e00);
        for (int ri = 0; ri < j.nrows; ++ri)
          for (int cbi = 0; cbi < j.CbG; ++cbi)
            for (int seg = 0; seg < j.nseg; ++seg) {
              int64_t src16 = ((int64_t)(j.row_mul * y + j.grp_row[g][ri]) * j.a_Cb + j.grp_cb0[g] + cbi) * j.a_slab16 +
                              j.seg_src16[seg] + x0;
              int64_t dst16 = (int64_t)(ri * j.CbG + cbi) * j.pslab16 + j.seg_dst16[seg];
              if (src16 < 0 || src16 + j.seg_len16[seg] > (int64_t)op.elems16) { set_error("src OOB"); return 4; }
              if (dst16 + j.seg_len16[seg] > j.stage16) { set_error("dst OOB"); return 5; }
              for (int64_t e = 0; e < (int64_t)j.seg_len16[seg] * 8; ++e) {
                st_hi[dst16 * 8 + e] = hi[src16 * 8 + e];
                st_lo[dst16 * 8 + e] = lo[src16 * 8 + e];
              }
            }
        if (j.rf_R) {
          // row-fold: patch row iy -> output rows r_min..r_max, weights slice of the descending-ky image (conv.cuh)
          const int KH = j.rf_kh, R = j.rf_R, nblk = j.rf_nblk, NR = KH * nblk;
          for (int ri = 0; ri < j.nrows; ++ri) {
            const int iy = g * j.nrows + ri;
            std::vector<double> &accw = (j.ksplit && (iy & 1)) ? acc2 : acc;  // K-split: patch rows by parity
            const int nf = (iy < R ? 1 : 0) + ((j.ksplit && iy >= 1 && iy - 1 < R) ? 1 : 0);
            const int r_min = std::max(0, iy - (KH - 1)), r_max = std::min(R - 1, iy);
            const int nb = r_max - r_min + 1, blk0 = KH - 1 - (iy - r_min);
            {  // the kernel reads these from the host-computed per-row table
              const uint32_t ib = (1u << 4) | ((uint32_t)(kTileM >> 4) << 24);
              auto idn = [&](int rows) { return rows > 0 ? ib | ((uint32_t)((rows * nblk) >> 3) << 17) : 0u; };
              if (j.rf_dcol[iy] != (uint32_t)(r_min * nblk) || j.rf_boff[iy] != (uint32_t)(blk0 * nblk) ||
                  j.rf_idn_all[iy] != idn(nb) || j.rf_idn_acc[iy] != idn(nb - nf) || j.rf_idn_new[iy] != idn(nf) ||
                  j.rf_off_new[iy] != (uint32_t)((nb - nf) * nblk)) { set_error("row-fold table mismatch at row %d", iy); return 11; }
            }
            for (int st = 0; st < j.rf_steps; ++st) {
              const KStep ks = j.steps[st];
              const uint16_t *chunk = pk.data() + (size_t)st * j.chunk16 * 8;   // [hi: 2 x NR rows][lo: 2 x NR rows]
              const bool fresh = (st == 0 && nf > 0);
              mma_count += 3 * (fresh && nb > nf ? 2 : 1);
              for (int m = 0; m < kTileM; ++m)
                for (int u = 0; u < 2; ++u) {
                  int64_t a16 = (int64_t)ri * j.rf_row16 + ks.a_off16 + (int64_t)u * ks.lbo16 + m;
                  if (a16 >= j.stage16) { set_error("A desc OOB (rowfold)"); return 6; }
                  for (int i = 0; i < 8; ++i) {
                    double ah = h2f_bits(st_hi[a16 * 8 + i]), al = h2f_bits(st_lo[a16 * 8 + i]);
                    for (int n = 0; n < nb * nblk; ++n) {
                      int64_t brow = (int64_t)blk0 * nblk + n;              // row inside one k-half image
                      if (brow >= NR) { set_error("B desc OOB (rowfold)"); return 6; }
                      double bh = h2f_bits(chunk[((int64_t)u * NR + brow) * 8 + i]);
                      double bl = h2f_bits(chunk[((int64_t)(2 + u) * NR + brow) * 8 + i]);
                      size_t col = (size_t)r_min * nblk + n;               // accumulator column
                      double &d = accw[(size_t)m * 512 + col];
                      const bool overwrite = fresh && (n >= (nb - nf) * nblk) && u == 0 && i == 0;
                      if (overwrite) d = 0;                                  // accumulate = 0 on the first MMA of the new row
                      d += ah * bh + al * bh + ah * bl;
                    }
                  }
                }
            }
          }
          continue;
        }
        if (j.pf) {
          // phase-fold: chunk = tap; N = pf_n[ch] columns starting at pf_col[ch]
          for (int ch = 0; ch < 4; ++ch) {
            const int n_ = j.pf_n[ch];
            if (j.pf_len16[ch] > j.chunk16) { set_error("pf chunk larger than slot"); return 6; }
            const uint16_t *chunk = pk.data() + ((size_t)g * j.pf_grp16 + j.pf_src16[ch]) * 8;
            const uint16_t *b_hi = chunk, *b_lo = chunk + (size_t)j.spc * 2 * n_ * 8;
            for (int st = 0; st < j.spc; ++st) {
              const KStep ks = j.steps[ch * j.spc + st];
              mma_count += 3;
              const int w = j.ksplit ? (st & 1) : 0;
              std::vector<double> &accw = w ? acc2 : acc;
              if (j.ksplit && first_w[w]) {  // accumulate = 0: must cover every column
                if (j.pf_col[ch] != 0 || n_ != Npad) { set_error("pf ksplit: first MMA does not cover all columns"); return 6; }
                for (int m = 0; m < kTileM; ++m) for (int n = 0; n < n_; ++n) accw[(size_t)m * Npad + n] = 0;
                first_w[w] = false;
              }
              for (int m = 0; m < kTileM; ++m)
                for (int u = 0; u < 2; ++u) {
                  int64_t a16 = (int64_t)ks.a_off16 + (int64_t)u * ks.lbo16 + m;
                  if (a16 >= j.stage16) { set_error("A desc OOB (pf)"); return 6; }
                  for (int i = 0; i < 8; ++i) {
                    double ah = h2f_bits(st_hi[a16 * 8 + i]), al = h2f_bits(st_lo[a16 * 8 + i]);
                    for (int n = 0; n < n_; ++n) {
                      int64_t b16 = (int64_t)st * 2 * n_ + (int64_t)u * n_ + n;
                      if ((b16 + 1) > j.pf_len16[ch] / 2) { set_error("B desc OOB (pf)"); return 6; }
                      double bh = h2f_bits(b_hi[b16 * 8 + i]), bl = h2f_bits(b_lo[b16 * 8 + i]);
                      accw[(size_t)m * Npad + j.pf_col[ch] + n] += ah * bh + al * bh + ah * bl;
                    }
                  }
                }
            }
          }
          continue;
        }
        for (int ch = 0; ch < j.nchunks; ++ch) {
          const uint16_t *chunk = pk.data() + (size_t)(g * j.nchunks + ch) * j.chunk16 * 8;
          const uint16_t *b_hi = chunk, *b_lo = chunk + (size_t)j.spc * 2 * Npad * 8;
          for (int st = 0; st < j.spc; ++st) {
            const KStep ks = j.steps[ch * j.spc + st];
            mma_count += 3 * j.mt;
            const int w = j.ksplit ? ((sc + st) & 1) : 0;
            std::vector<double> &accw = w ? acc2 : acc;
            if (j.ksplit && first_w[w]) { std::fill(accw.begin(), accw.end(), 0.0); first_w[w] = false; }
            for (int t = 0; t < j.mt; ++t)
            for (int m = 0; m < kTileM; ++m)
              for (int u = 0; u < 2; ++u) {
                // row m: +16 B (SBO = 8 rows * 16 B); second output row of an mt=2 unit: one patch row lower
                int64_t a16 = (int64_t)ks.a_off16 + (int64_t)u * ks.lbo16 + m + (int64_t)t * j.CbG * j.pslab16;
                if (a16 >= j.stage16) { set_error("A desc OOB"); return 6; }
                for (int i = 0; i < 8; ++i) {
                  double ah = h2f_bits(st_hi[a16 * 8 + i]), al = h2f_bits(st_lo[a16 * 8 + i]);
                  for (int n = 0; n < Npad; ++n) {
                    int64_t b16 = (int64_t)st * 2 * Npad + (int64_t)u * Npad + n;
                    double bh = h2f_bits(b_hi[b16 * 8 + i]), bl = h2f_bits(b_lo[b16 * 8 + i]);
                    accw[((size_t)t * kTileM + m) * Npad + n] += ah * bh + al * bh + ah * bl;
                  }
                }
              }
          }
          sc += j.spc;
        }
      }
      // epilogue: K-split units add the second warp's accumulator
      if (j.ksplit)
        for (size_t i = 0; i < acc.size(); ++i) acc[i] += acc2[i];
      // epilogue placement
      if (j.rf_R) {
        const int nblk = j.rf_nblk;
        for (int t = 0; t < j.rf_R; ++t)
          for (int m = 0; m < (j.xfold_kw ? j.tile_dx : kTileM); ++m) {
            int xx = x0 + m;
            if (xx >= j.Wo || y + t >= j.Ho) continue;
            written[(size_t)(y + t) * Wo + xx]++;
            for (int n = 0; n < cout; ++n) {
              double v = 0;
              if (j.xfold_kw) for (int kx = 0; kx < j.xfold_kw; ++kx) v += acc[(size_t)(m + kx) * 512 + (size_t)t * nblk + kx * cout + n];
              else v = acc[(size_t)m * 512 + (size_t)t * nblk + n];
              out[((size_t)n * Ho + y + t) * Wo + xx] = v;
            }
          }
        continue;
      }
      if (j.pf) {
        for (int m = 0; m < kTileM; ++m) {
          int xx = x0 + m;
          if (xx >= j.Wo || y >= j.Ho) continue;
          for (int blk = 0; blk < 4; ++blk) {
            const int a = blk >> 1, b = (blk == 1 || blk == 2) ? 1 : 0;
            int yo = 2 * y + a, xo = 2 * xx + b;
            if (yo >= Ho || xo >= Wo) { set_error("output OOB (pf)"); return 7; }
            written[(size_t)yo * Wo + xo]++;
            for (int n = 0; n < cout; ++n) out[((size_t)n * Ho + yo) * Wo + xo] = acc[(size_t)m * Npad + blk * j.pf_cout + n];
          }
        }
        continue;
      }
      if (j.xfold_kw) {
        for (int m = 0; m < j.tile_dx; ++m) {
          int xx = x0 + m;
          if (xx >= j.Wo) continue;
          written[(size_t)y * Wo + xx]++;
          for (int n = 0; n < cout; ++n) {
            double sacc = 0;
            for (int kx = 0; kx < j.xfold_kw; ++kx) sacc += acc[(size_t)(m + kx) * Npad + kx * cout + n];
            out[((size_t)n * Ho + y) * Wo + xx] = sacc;
          }
        }
        continue;
      }
      for (int t = 0; t < j.mt; ++t)
      for (int m = 0; m < kTileM; ++m) {
        int xx = x0 + m;
        if (xx >= j.Wo || y + t >= j.Ho) continue;
        int yo = (y + t) * j.oy_mul + j.oy_off, xo = xx * j.ox_mul + j.ox_off;
        if (yo >= Ho || xo >= Wo) { set_error("output OOB"); return 7; }
        written[(size_t)yo * Wo + xo]++;
        for (int n = 0; n < cout; ++n) out[((size_t)n * Ho + yo) * Wo + xo] = acc[((size_t)t * kTileM + m) * Npad + n];
      }
    }
  }
  for (char wv : written)
    if (wv != 1) { set_error("output pixel written %d times", (int)wv); return 8; }
  // direct reference
  if (!transposed) {
    for (int co = 0; co < cout; ++co)
      for (int oy = 0; oy < Ho; ++oy)
        for (int ox = 0; ox < Wo; ++ox) {
          double s = 0;
          for (int ci = 0; ci < cin; ++ci)
            for (int ky = 0; ky < k; ++ky) {
              int iy = oy * stride + ky - pad;
              if (iy < 0 || iy >= H) continue;
              for (int kx = 0; kx < k; ++kx) {
                int ix = ox * stride + kx - pad;
                if (ix < 0 || ix >= W) continue;
                s += (double)w[(((size_t)co * cin + ci) * k + ky) * k + kx] * xq[((size_t)ci * H + iy) * W + ix];
              }
            }
          ref[((size_t)co * Ho + oy) * Wo + ox] = s;
        }
  } else {
    for (int ci = 0; ci < cin; ++ci)
      for (int iy = 0; iy < H; ++iy)
        for (int ix = 0; ix < W; ++ix) {
          double xv = xq[((size_t)ci * H + iy) * W + ix];
          for (int ky = 0; ky < k; ++ky) {
            int oy = iy * stride - pad + ky;
            if (oy < 0 || oy >= Ho) continue;
            for (int kx = 0; kx < k; ++kx) {
              int ox = ix * stride - pad + kx;
              if (ox < 0 || ox >= Wo) continue;
              for (int co = 0; co < cout; ++co)
                ref[((size_t)co * Ho + oy) * Wo + ox] += xv * w[(((size_t)ci * cout + co) * k + ky) * k + kx];
            }
          }
        }
  }
  double me = 0, mr = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    double d = std::fabs(out[i] - ref[i]);
    if (!(d <= 1e300)) d = 1e300;  // NaN -> huge
    me = std::max(me, d);
    mr = std::max(mr, std::fabs(ref[i]));
  }
  *max_err = me; *max_ref = mr;
  if (smem_bytes) *smem_bytes = (int)smem_max;
  if (n_mma) *n_mma = mma_count;
  return 0;
}

// planner decisions for one layer (first job): {mt, ksplit, pf, rf_R, xfold_kw, b_resident, a_stages, b_slots, ntiles, Npad}
extern "C" int emu_plan_info(int cin, int cout, int k, int stride, int pad, int transposed, int adj, int H, int W, int *out10) {
  ConvDef c;
  init_conv_def(c, "emu", cin, cout, k, stride, pad, transposed != 0, adj);
  build_phases(c);
  Operand op = operand_geometry(cin, H, W, &c);
  ConvPhase &ph = c.has_fold ? c.fold : c.phases[0];
  if (!ph.pf && build_phase_tables(c, ph) != FAV_OK) return 1;
  ConvJob j;
  if (fill_conv_job(c, ph, op, j) != FAV_OK) return 2;
  conv_tc_choose_slots(j);
  const int v[10] = {j.mt, j.ksplit, j.pf, j.rf_R, j.xfold_kw, j.b_resident, j.a_stages, j.b_slots, j.ntiles, j.Npad};
  for (int i = 0; i < 10; ++i) out10[i] = v[i];
  return 0;
}


extern "C" int emu_plan_detail(int cin, int cout, int k, int stride, int pad, int transposed, int adj, int H, int W, int *out16) {
  ConvDef c;
  init_conv_def(c, "emu", cin, cout, k, stride, pad, transposed != 0, adj);
  build_phases(c);
  Operand op = operand_geometry(cin, H, W, &c);
  ConvPhase &ph = c.has_fold ? c.fold : c.phases[0];
  if (!ph.pf && build_phase_tables(c, ph) != FAV_OK) return 1;
  ConvJob j;
  if (fill_conv_job(c, ph, op, j) != FAV_OK) return 2;
  conv_tc_choose_slots(j);
  const int v[16] = {j.ngroups, j.nchunks, j.spc, j.nrows, j.CbG, j.nseg, j.a_stages, j.b_slots, j.b_resident, j.stage16 * 16,
                     j.chunk16 * 16, j.pslab16, j.tiles_x, j.ntiles, j.Npad, j.ksplit};
  for (int i = 0; i < 16; ++i) out16[i] = v[i];
  return 0;
}

// ---- conv_res.cu (swapped roles: weights = M operand, pixels = N operand, cost-balanced tile table) ---------------------
// Emulates the kernel's addressing from the same host planner (conv_res_plan.hpp): bulk-copy sources, stage layout, the two
// matrix descriptors per K step, accumulator rows / columns, planar output placement; checks it against a direct convolution
// and that every output pixel is produced exactly once.  info4 = {grid, tiles, max CTA cost, sum of CTA costs}.
extern "C" int emu_res_check(int cin, int pad, int H, int W, int nctas, unsigned seed, double *max_err, double *max_ref, int *info4) {
  const int cout = 128, k = 3;
  ConvDef c;
  init_conv_def(c, "emu_res", cin, cout, k, 1, pad, false, 0);
  build_phases(c);
  ConvPhase &ph = c.phases[0];
  if (build_phase_tables(c, ph) != FAV_OK) return 1;
  if (!conv_res_eligible(c, ph)) { set_error("layer not eligible for the residual kernel"); return 2; }
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  std::vector<float> w((size_t)cin * cout * k * k), x((size_t)cin * H * W);
  for (auto &v : w) v = U(rng) * 0.1f;
  for (auto &v : x) v = U(rng) * 3.f;
  Operand op = operand_geometry(cin, H, W, &c);
  std::vector<uint16_t> hi(op.elems16 * 8, 0), lo(op.elems16 * 8, 0);
  std::vector<float> xq((size_t)cin * H * W);
  for (int ci = 0; ci < cin; ++ci)
    for (int y = 0; y < H; ++y)
      for (int xx = 0; xx < W; ++xx) {
        float v = x[((size_t)ci * H + y) * W + xx];
        uint16_t hb = f2h_bits(v), lb = f2h_bits(v - h2f_bits(hb));
        int64_t o = op.off16(op.padT + y, ci / 8, op.padL + xx) * 8 + ci % 8;
        hi[o] = hb; lo[o] = lb;
        xq[((size_t)ci * H + y) * W + xx] = h2f_bits(hb) + h2f_bits(lb);
      }
  std::vector<uint16_t> pk = pack_phase_weights(c, ph, w);
  ResJob j;
  if (fill_res_job(c, ph, op, j) != FAV_OK) return 3;
  const int Ho = j.Ho, Wo = j.Wo;
  ResPlan pl = plan_res_tiles(Ho, Wo, nctas);
  if ((int)pl.cta_first.size() != pl.grid + 1 || pl.cta_first.back() != (int)pl.tiles.size()) { set_error("tile table shape"); return 4; }
  const int Wp = round_up(Wo, 16);
  std::vector<double> out((size_t)cout * Ho * Wp, 0.0);
  std::vector<int> written((size_t)Ho * Wo, 0);
  const size_t stage16 = (size_t)4 * kResCbG * kResPslab;  // one plane
  const size_t chunk16 = (size_t)2 * kResSpc * 2 * 128;
  std::vector<uint16_t> st_hi(stage16 * 8), st_lo(stage16 * 8);
  std::vector<double> acc((size_t)2 * 128 * kResMaxNt);
  for (int cta = 0; cta < pl.grid; ++cta) {
    if (pl.cta_first[cta + 1] <= pl.cta_first[cta]) { set_error("CTA %d has no tile", cta); return 5; }
    for (int ti = pl.cta_first[cta]; ti < pl.cta_first[cta + 1]; ++ti) {
      const ResTile t = pl.tiles[ti];
      if (t.nt <= 0 || t.nt > kResMaxNt || t.nt % 16 || t.x0 % 16 || t.y % 2 || t.y >= Ho || t.x0 >= Wo) { set_error("bad tile"); return 6; }
      std::fill(acc.begin(), acc.end(), 1e30);  // stale TMEM must be overwritten by the first MMA (accumulate = 0)
      bool first = true;
      for (int g = 0; g < j.ngroups; ++g) {
        std::fill(st_hi.begin(), st_hi.end(), (uint16_t)0x7e00);  // NaN poison: reading an unloaded byte is a bug
        std::fill(st_lo.begin(), st_lo.end(), (uint16_t)0x7e00);
        for (int lane = 0; lane < 4 * kResCbG * 2; ++lane) {  // the producer warp: lane = (patch row, channel block, hi/lo)
          const int ri = lane / (2 * kResCbG), cbi = (lane >> 1) % kResCbG, part = lane & 1;
          const int64_t src16 = ((int64_t)(t.y + ri + j.in_row0) * j.a_Cb + g * kResCbG + cbi) * j.a_slab16 + t.x0 + j.in_col0;
          const int64_t dst16 = (int64_t)(ri * kResCbG + cbi) * kResPslab, len16 = t.nt + 2;
          if (src16 < 0 || src16 + len16 > (int64_t)op.elems16) { set_error("src OOB"); return 7; }
          if (dst16 + len16 > (int64_t)stage16) { set_error("dst OOB"); return 8; }
          std::vector<uint16_t> &dstv = part ? st_lo : st_hi;
          const std::vector<uint16_t> &srcv = part ? lo : hi;
          for (int64_t e = 0; e < len16 * 8; ++e) dstv[dst16 * 8 + e] = srcv[src16 * 8 + e];
        }
        int sidx = 0;
        for (int ch = 0; ch < kResChunks; ++ch) {
          const uint16_t *chunk = pk.data() + ((size_t)g * kResChunks + ch) * chunk16 * 8;
          for (int st = 0; st < kResSpc; ++st, ++sidx) {
            const uint32_t dls = j.steps[sidx];
            const int a_off16 = dls & 0xffff, lbo16 = dls >> 16;
            for (int r = 0; r < 2; ++r)
              for (int u = 0; u < 2; ++u)
                for (int m = 0; m < 128; ++m) {
                  const int64_t w16 = (int64_t)st * 256 + (int64_t)u * 128 + m;  // LBO = 128 rows, SBO: 8-row groups contiguous
                  const uint16_t *wh = chunk + w16 * 8, *wl = chunk + ((int64_t)kResSpc * 256 + w16) * 8;
                  for (int n = 0; n < t.nt; ++n) {
                    const int64_t p16 = (int64_t)r * kResCbG * kResPslab + a_off16 + (int64_t)u * lbo16 + n;
                    if (p16 >= (int64_t)stage16) { set_error("patch descriptor OOB"); return 9; }
                    double d = (first && u == 0) ? 0.0 : acc[((size_t)r * 128 + m) * kResMaxNt + n];
                    for (int i = 0; i < 8; ++i) {
                      const double a_h = h2f_bits(wh[i]), a_l = h2f_bits(wl[i]);
                      const double b_h = h2f_bits(st_hi[p16 * 8 + i]), b_l = h2f_bits(st_lo[p16 * 8 + i]);
                      d += a_h * b_h + a_l * b_h + a_h * b_l;
                    }
                    acc[((size_t)r * 128 + m) * kResMaxNt + n] = d;
                  }
                }
            first = false;
          }
        }
      }
      for (int r = 0; r < 2; ++r) {
        const int yo = t.y + r;
        if (yo >= Ho) continue;
        for (int n = 0; n < t.nt; ++n) {
          if (t.x0 + n >= Wp) { set_error("store beyond the row pitch"); return 10; }
          if (t.x0 + n < Wo) written[(size_t)yo * Wo + t.x0 + n]++;
          for (int m = 0; m < cout; ++m) out[((size_t)m * Ho + yo) * Wp + t.x0 + n] = acc[((size_t)r * 128 + m) * kResMaxNt + n];
        }
      }
    }
  }
  for (int wv : written)
    if (wv != 1) { set_error("output pixel written %d times", wv); return 11; }
  double me = 0, mr = 0;
  for (int co = 0; co < cout; ++co)
    for (int oy = 0; oy < Ho; ++oy)
      for (int ox = 0; ox < Wo; ++ox) {
        double sref = 0;
        for (int ci = 0; ci < cin; ++ci)
          for (int ky = 0; ky < k; ++ky) {
            int iy = oy + ky - pad;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < k; ++kx) {
              int ix = ox + kx - pad;
              if (ix < 0 || ix >= W) continue;
              sref += (double)w[(((size_t)co * cin + ci) * k + ky) * k + kx] * xq[((size_t)ci * H + iy) * W + ix];
            }
          }
        double dv = std::fabs(out[((size_t)co * Ho + oy) * Wp + ox] - sref);
        if (!(dv <= 1e300)) dv = 1e300;
        me = std::max(me, dv); mr = std::max(mr, std::fabs(sref));
      }
  *max_err = me; *max_ref = mr;
  if (info4) { info4[0] = pl.grid; info4[1] = (int)pl.tiles.size(); info4[2] = pl.max_cost; info4[3] = pl.sum_cost; }
  return 0;
}

// tile-table statistics only (full-size layers): info6 = {grid, tiles, max cost, sum cost, narrowest tile, widest tile}
extern "C" int emu_res_plan(int Ho, int Wo, int nctas, int *info6) {
  ResPlan pl = plan_res_tiles(Ho, Wo, nctas);
  std::vector<int> cover((size_t)((Ho + 1) / 2) * ((Wo + 15) / 16), 0);
  int mn = 1 << 30, mx = 0;
  for (const ResTile &t : pl.tiles) {
    mn = std::min(mn, (int)t.nt); mx = std::max(mx, (int)t.nt);
    for (int g = 0; g < t.nt / 16; ++g) cover[(size_t)(t.y / 2) * ((Wo + 15) / 16) + t.x0 / 16 + g]++;
  }
  for (int v : cover)
    if (v != 1) { set_error("granule covered %d times", v); return 1; }
  info6[0] = pl.grid; info6[1] = (int)pl.tiles.size(); info6[2] = pl.max_cost; info6[3] = pl.sum_cost; info6[4] = mn; info6[5] = mx;
  return 0;
}
