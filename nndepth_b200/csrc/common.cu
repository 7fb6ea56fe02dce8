// Error plumbing and ABI bookkeeping of libnndepth_b200 (no kernels here).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace nnd {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

nnd_status cuda_fail(cudaError_t e, const char* where) {
  set_error("%s: CUDA error %d (%s)", where, static_cast<int>(e), cudaGetErrorString(e));
  return NND_ERR_CUDA;
}

nnd_status check_launch(const char* kernel) {
  // cudaPeekAtLastError does not synchronise, so this stays CUDA-graph capturable.
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky launch-configuration error
    return cuda_fail(e, kernel);
  }
  return NND_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached_sms = n;
    cached_dev = dev;
  }
  return cached_sms;
}

}  // namespace nnd

extern "C" {

int nnd_abi_version(void) { return NND_ABI_VERSION; }

const char* nnd_last_error_string(void) { return nnd::g_error; }

int nnd_row_pitch(int width) { return width <= 0 ? 0 : (width + 3) & ~3; }

}  // extern "C"
