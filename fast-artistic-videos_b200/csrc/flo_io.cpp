// flo_io.cpp -- a-3 / a-13: Middlebury .flo on the host.
//   flowFileLoader_load           flowFileLoader.lua:17-37   (tag read but NOT validated :20; output [dy,dx])
//   readMiddlebury                consistencyChecker/consistencyChecker.cpp:16-36 (output planes u, v)
// The reference parses the payload in an interpreted Lua double loop over H*W (flowFileLoader.lua:28-34); here it is
// one fread plus a strided de-interleave.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../include/fav.h"

namespace fav { void set_error(const char *fmt, ...); }

extern "C" {

int fav_flo_read_header(const char *path, int *W, int *H) {
  if (!path || !W || !H) { fav::set_error("fav_flo_read_header: null argument"); return FAV_ERR_INVALID; }
  FILE *f = fopen(path, "rb");
  if (!f) { fav::set_error("Could not open %s", path); return FAV_ERR_IO; }
  float tag; int w, h;
  bool ok = fread(&tag, 4, 1, f) == 1 && fread(&w, 4, 1, f) == 1 && fread(&h, 4, 1, f) == 1;
  fclose(f);
  if (!ok || w <= 0 || h <= 0) { fav::set_error("%s: truncated or invalid .flo header", path); return FAV_ERR_IO; }
  *W = w; *H = h;
  return FAV_OK;
}

int fav_flo_read(const char *path, float *out, int layout) {
  if (!path || !out) { fav::set_error("fav_flo_read: null argument"); return FAV_ERR_INVALID; }
  FILE *f = fopen(path, "rb");
  if (!f) { fav::set_error("Could not open %s", path); return FAV_ERR_IO; }
  float tag; int W, H;
  if (!(fread(&tag, 4, 1, f) == 1 && fread(&W, 4, 1, f) == 1 && fread(&H, 4, 1, f) == 1) || W <= 0 || H <= 0) {
    fclose(f);
    fav::set_error("%s: truncated or invalid .flo header", path);
    return FAV_ERR_IO;
  }
  const size_t n = (size_t)W * H;
  std::vector<float> raw(2 * n);
  size_t got = fread(raw.data(), sizeof(float), 2 * n, f);
  fclose(f);
  if (got != 2 * n) { fav::set_error("%s: truncated .flo payload", path); return FAV_ERR_IO; }
  float *p0 = out, *p1 = out + n;  // layout 0: [dy(v), dx(u)]  (flowFileLoader.lua:31-32); layout 1: [u, v]
  if (layout == 0) { p0 = out + n; p1 = out; }
  for (size_t i = 0; i < n; ++i) { p0[i] = raw[2 * i]; p1[i] = raw[2 * i + 1]; }
  return FAV_OK;
}
}
