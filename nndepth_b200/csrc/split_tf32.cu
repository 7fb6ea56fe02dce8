// Error-compensated TF32 ("3xTF32") operand split for the recurrent convolutions (sm_100a).
//
// TF32 tensor-core convolutions in the ConvGRU recurrence move the final disparity 0.013 px from the fp32
// reference (bar: 0.01 px), fp32 CUDA-core convolutions cost 4x the time.  Splitting every operand into a
// TF32-exact head and an fp32 tail,  x = hi + lo,  hi = RN_tf32(x),  lo = x - hi  (exact in fp32),
//   conv(x, w) = conv(hi, w_hi) + conv(lo, w_hi) + conv(hi, w_lo) + O(2^-22)
// keeps fp32-level accuracy on the tensor cores.  The three products are ONE convolution over 3x the
// input channels: activations [hi ; lo ; hi] against weights [w_hi ; w_hi ; w_lo].
//
//   nnd_split_tf32: x (N, C, H*W) -> out (N, 3C, H*W) = [hi ; lo ; hi], one pass (read 1x, write 3x).
#include "common.cuh"

namespace nnd {

__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}

// chw4 = C * H * W / 4 float4 per image
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float4* __restrict__ x, long long chw4, long long total4, float4* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / chw4, r = i - n * chw4;
    const float4 v = __ldg(x + i);
    float4 hi, lo;
    hi.x = rn_tf32(v.x); hi.y = rn_tf32(v.y); hi.z = rn_tf32(v.z); hi.w = rn_tf32(v.w);
    lo.x = __fsub_rn(v.x, hi.x); lo.y = __fsub_rn(v.y, hi.y); lo.z = __fsub_rn(v.z, hi.z); lo.w = __fsub_rn(v.w, hi.w);
    float4* o = out + n * 3 * chw4 + r;
    o[0] = hi;
    o[chw4] = lo;
    o[2 * chw4] = hi;
  }
}

__global__ void __launch_bounds__(256)
split_tf32_scalar_kernel(const float* __restrict__ x, long long chw, long long total, float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / chw, r = i - n * chw;
    const float v = __ldg(x + i);
    const float hi = rn_tf32(v);
    float* o = out + n * 3 * chw + r;
    o[0] = hi;
    o[chw] = __fsub_rn(v, hi);
    o[2 * chw] = hi;
  }
}

}  // namespace nnd

extern "C" {

nnd_status nnd_split_tf32(const float* x, int N, int C, long long hw, float* out, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NND_REQUIRE(x && out, "split_tf32: null pointer");
  NND_REQUIRE(N > 0 && C > 0 && hw > 0, "split_tf32: N, C, H*W must be positive");
  const long long chw = static_cast<long long>(C) * hw;
  const long long total = chw * N;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (chw % 4 == 0 && aligned16(x) && aligned16(out)) {
    const long long want = (total / 4 + 255) / 256;
    split_tf32_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(x), chw / 4, total / 4, reinterpret_cast<float4*>(out));
  } else {
    const long long want = (total + 255) / 256;
    split_tf32_scalar_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, stream>>>(x, chw, total, out);
  }
  return check_launch("split_tf32_kernel");
}

}  // extern "C"
