// Tensor-core (tcgen05 / TMEM / TMA) build of the RAFT-Stereo correlation pyramid -- placeholder
// until the kernel lands; the entry point reports NND_ERR_UNSUPPORTED instead of silently falling
// back to another precision.
#include "common.cuh"

namespace nnd {

nnd_status corr1d_build_tf32(const float*, const float*, int, int, int, int, int, int, float* const*, const int*,
                             cudaStream_t) {
  set_error("corr1d_build: the TF32 tensor-core path is not built into this library");
  return NND_ERR_UNSUPPORTED;
}

}  // namespace nnd
