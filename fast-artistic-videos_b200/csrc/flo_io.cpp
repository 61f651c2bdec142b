// flo_io.cpp -- a-3 / a-13 / f-2: the file formats of the pipeline, host side.
//   flowFileLoader_load           flowFileLoader.lua:17-37   (tag read but NOT validated :20; output [dy,dx])
//   readMiddlebury                consistencyChecker/consistencyChecker.cpp:16-36 (output planes u, v)
//   image.load(ppm|pgm)           fast_artistic_video.lua:95,103 (byte / 255 as float; third-party `image` rock)
//   CTensor::readFromPPM          consistencyChecker/CTensor.h:888-936 (comment lines, P5/P6)
// The reference parses a .flo payload in an interpreted Lua double loop over H*W (flowFileLoader.lua:28-34); here it is
// one fread plus a strided de-interleave straight into the caller's (pinned) buffer.  Nothing in this file throws across the
// C ABI: header fields are validated against the file size before anything is sized from them.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <memory>
#include <new>

#include "../../include/fav.h"

namespace fav { void set_error(const char *fmt, ...); }

namespace {
struct File {
  FILE *f = nullptr;
  explicit File(const char *path) : f(fopen(path, "rb")) {}
  ~File() { if (f) fclose(f); }
};

long long file_size(FILE *f) {
  struct stat st;
  if (fstat(fileno(f), &st) != 0) return -1;
  return (long long)st.st_size;
}

// header of a .flo: tag (ignored, flowFileLoader.lua:20), int32 W, int32 H; payload must be present in full
int flo_header(FILE *f, const char *path, int *W, int *H) {
  float tag; int w, h;
  if (!(fread(&tag, 4, 1, f) == 1 && fread(&w, 4, 1, f) == 1 && fread(&h, 4, 1, f) == 1) || w <= 0 || h <= 0 ||
      (long long)w * h > (1ll << 34)) {
    fav::set_error("%s: truncated or invalid .flo header", path);
    return FAV_ERR_IO;
  }
  const long long need = 12 + 8ll * w * h, have = file_size(f);
  if (have >= 0 && have < need) { fav::set_error("%s: truncated .flo payload (%lld of %lld bytes)", path, have, need); return FAV_ERR_IO; }
  *W = w; *H = h;
  return FAV_OK;
}

// binary PNM header "P5|P6 <W> <H> <maxval>\n" with '#' comment lines (CTensor.h:899-915); leaves f at the first payload byte
int pnm_header(FILE *f, const char *path, int *W, int *H, int *C) {
  int c0 = fgetc(f), c1 = fgetc(f);
  if (c0 != 'P' || (c1 != '5' && c1 != '6')) { fav::set_error("%s: not a binary PGM/PPM (P5/P6)", path); return FAV_ERR_IO; }
  int vals[3], n = 0;
  while (n < 3) {
    int ch = fgetc(f);
    if (ch == EOF) { fav::set_error("%s: truncated PNM header", path); return FAV_ERR_IO; }
    if (ch == '#') { while (ch != '\n' && ch != EOF) ch = fgetc(f); continue; }
    if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') continue;
    if (ch < '0' || ch > '9') { fav::set_error("%s: bad PNM header", path); return FAV_ERR_IO; }
    long long v = 0;
    while (ch >= '0' && ch <= '9') { v = v * 10 + (ch - '0'); if (v > (1 << 30)) break; ch = fgetc(f); }
    vals[n++] = (int)v;  // the single whitespace byte after the number is consumed by the loop above
  }
  if (vals[0] <= 0 || vals[1] <= 0 || vals[2] != 255) { fav::set_error("%s: unsupported PNM header (8-bit only)", path); return FAV_ERR_IO; }
  *W = vals[0]; *H = vals[1]; *C = c1 == '6' ? 3 : 1;
  return FAV_OK;
}
}  // namespace

extern "C" {

int fav_flo_read_header(const char *path, int *W, int *H) {
  if (!path || !W || !H) { fav::set_error("fav_flo_read_header: null argument"); return FAV_ERR_INVALID; }
  File fl(path);
  if (!fl.f) { fav::set_error("Could not open %s", path); return FAV_ERR_IO; }
  return flo_header(fl.f, path, W, H);
}

// out: host [2,H,W]; capacity_floats = number of floats `out` can hold (the file is rejected if it needs more: the header
// is re-read here, and a producer may have rewritten the file since fav_flo_read_header)
int fav_flo_read(const char *path, float *out, size_t capacity_floats, int layout) {
  if (!path || !out) { fav::set_error("fav_flo_read: null argument"); return FAV_ERR_INVALID; }
  File fl(path);
  if (!fl.f) { fav::set_error("Could not open %s", path); return FAV_ERR_IO; }
  int W, H;
  int rc = flo_header(fl.f, path, &W, &H);
  if (rc != FAV_OK) return rc;
  const size_t n = (size_t)W * H;
  if (2 * n > capacity_floats) {
    fav::set_error("%s: %dx%d flow does not fit the caller's buffer (%zu floats)", path, W, H, capacity_floats);
    return FAV_ERR_IO;
  }
  std::unique_ptr<float[]> raw(new (std::nothrow) float[2 * n]);
  if (!raw) { fav::set_error("%s: out of memory for a %dx%d flow", path, W, H); return FAV_ERR_IO; }
  if (fread(raw.get(), sizeof(float), 2 * n, fl.f) != 2 * n) { fav::set_error("%s: truncated .flo payload", path); return FAV_ERR_IO; }
  float *p0 = out, *p1 = out + n;  // layout 0: [dy(v), dx(u)]  (flowFileLoader.lua:31-32); layout 1: [u, v]
  if (layout == 0) { p0 = out + n; p1 = out; }
  const float *r = raw.get();
  for (size_t i = 0; i < n; ++i) { p0[i] = r[2 * i]; p1[i] = r[2 * i + 1]; }
  return FAV_OK;
}

int fav_pnm_read_header(const char *path, int *W, int *H, int *C) {
  if (!path || !W || !H || !C) { fav::set_error("fav_pnm_read_header: null argument"); return FAV_ERR_INVALID; }
  File fl(path);
  if (!fl.f) { fav::set_error("Could not open %s", path); return FAV_ERR_IO; }
  return pnm_header(fl.f, path, W, H, C);
}

// binary P6 / P5 -> planar fp32 [C,H,W] = byte / divisor (255: image.load's [0,1]; 1: readFromPPM's 0..255 planes)
int fav_pnm_read_f32(const char *path, float *out, size_t capacity_floats, float divisor) {
  if (!path || !out || !(divisor > 0)) { fav::set_error("fav_pnm_read_f32: bad argument"); return FAV_ERR_INVALID; }
  File fl(path);
  if (!fl.f) { fav::set_error("Could not open %s", path); return FAV_ERR_IO; }
  int W, H, C;
  int rc = pnm_header(fl.f, path, &W, &H, &C);
  if (rc != FAV_OK) return rc;
  const size_t n = (size_t)W * H;
  if (n * C > capacity_floats) { fav::set_error("%s: %dx%dx%d image does not fit the caller's buffer", path, W, H, C); return FAV_ERR_IO; }
  std::unique_ptr<unsigned char[]> raw(new (std::nothrow) unsigned char[n * C]);
  if (!raw) { fav::set_error("%s: out of memory", path); return FAV_ERR_IO; }
  if (fread(raw.get(), 1, n * C, fl.f) != n * C) { fav::set_error("%s: truncated PNM payload", path); return FAV_ERR_IO; }
  const unsigned char *r = raw.get();
  float lut[256];
  for (int v = 0; v < 256; ++v) lut[v] = (float)v / divisor;  // a correctly rounded fp32 division per byte value
  for (int c = 0; c < C; ++c) {
    float *dst = out + (size_t)c * n;
    for (size_t i = 0; i < n; ++i) dst[i] = lut[r[i * C + c]];
  }
  return FAV_OK;
}

// the raw payloads, for callers that convert on the GPU (fav_session_run_frame_bytes): the P5 / P6 bytes as stored
// (interleaved RGB) and the .flo (u,v) pairs as stored; sizes are returned and validated against the capacity
int fav_pnm_read_u8(const char *path, unsigned char *out, size_t capacity_bytes, int *W, int *H, int *C) {
  if (!path || !out || !W || !H || !C) { fav::set_error("fav_pnm_read_u8: null argument"); return FAV_ERR_INVALID; }
  File fl(path);
  if (!fl.f) { fav::set_error("Could not open %s", path); return FAV_ERR_IO; }
  int rc = pnm_header(fl.f, path, W, H, C);
  if (rc != FAV_OK) return rc;
  const size_t n = (size_t)*W * *H * *C;
  if (n > capacity_bytes) { fav::set_error("%s: %dx%dx%d image does not fit the caller's buffer", path, *W, *H, *C); return FAV_ERR_IO; }
  if (fread(out, 1, n, fl.f) != n) { fav::set_error("%s: truncated PNM payload", path); return FAV_ERR_IO; }
  return FAV_OK;
}

int fav_flo_read_raw(const char *path, float *out_uv_pairs, size_t capacity_floats, int *W, int *H) {
  if (!path || !out_uv_pairs || !W || !H) { fav::set_error("fav_flo_read_raw: null argument"); return FAV_ERR_INVALID; }
  File fl(path);
  if (!fl.f) { fav::set_error("Could not open %s", path); return FAV_ERR_IO; }
  int rc = flo_header(fl.f, path, W, H);
  if (rc != FAV_OK) return rc;
  const size_t n = 2 * (size_t)*W * *H;
  if (n > capacity_floats) { fav::set_error("%s: %dx%d flow does not fit the caller's buffer (%zu floats)", path, *W, *H, capacity_floats); return FAV_ERR_IO; }
  if (fread(out_uv_pairs, sizeof(float), n, fl.f) != n) { fav::set_error("%s: truncated .flo payload", path); return FAV_ERR_IO; }
  return FAV_OK;
}
}
