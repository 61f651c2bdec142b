// common.cu -- error state, launch counter, device probing.
#include "fav_common.cuh"

namespace fav {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    set_error("libfav_b200: no CUDA device available (%s); this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return FAV_ERR_NO_DEVICE;
  }
  return FAV_OK;
}
}  // namespace fav

extern "C" {
const char *fav_last_error(void) { return fav::g_err; }
int fav_version(void) { return 100; }
uint64_t fav_launch_count(void) { return fav::g_launches.load(); }
int fav_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}
}
