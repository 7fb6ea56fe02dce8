// Elementwise glue of the separable ConvGRU around its (cuDNN, TF32 tensor-core) convolutions, fused and
// channels-last (sm_100a).  Reference: nndepth/blocks/gru.py:5-37
//     hx = cat[h, x];  z = sigmoid(convz(hx));  r = sigmoid(convr(hx));
//     q = tanh(convq(cat[r*h, x]));  h = (1 - z) * h + z * q            (twice: 1x5 then 5x1 kernels)
//
// The convolutions run as error-compensated 3xTF32 (split_tf32.cu): their input is the operand split
// [hi ; lo ; hi] of cat[h, x].  Unfused, every half-step makes ~15 passes over tensors of up to 276 MB
// (cat, split, cuDNN's NCHW->NHWC conversion of the 1152-channel input, bias add, sigmoid, chunk, mul, cat,
// split, conversion, bias add, tanh, rsub, mul, mul, add).  Here one channels-last staging buffer
//     S[n][p][ h_hi(Ch) | h_lo(Ch) | h_hi(Ch) | x_hi(Cx) | x_lo(Cx) | x_hi(Cx) ]
// feeds all four convolutions of an iteration without any layout conversion (cuDNN's tensor-core kernels
// are NHWC-native): the x part is written once per iteration (the context half of x once per forward), the
// h part is overwritten in place -- by r*h for the q convolution, by the new h for the next half-step.
//
//   nnd_gru_stage    NCHW tensor -> [hi|lo|hi] at three channel offsets of S   (transpose + split)
//   nnd_gru_gate_r   zr_pre (NHWC, 2Ch) + bias, h -> z ; S.h <- split(r * h)
//   nnd_gru_gate_h   q_pre (NHWC, Ch) + bias, z, h -> h' = (1-z) h + z tanh(q) ; S.h <- split(h') ; h <- h'
#include "common.cuh"

namespace nnd {

__device__ __forceinline__ float rn_tf32_g(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = rn_tf32_g(v.x); hi.y = rn_tf32_g(v.y); hi.z = rn_tf32_g(v.z); hi.w = rn_tf32_g(v.w);
  lo.x = __fsub_rn(v.x, hi.x); lo.y = __fsub_rn(v.y, hi.y); lo.z = __fsub_rn(v.z, hi.z); lo.w = __fsub_rn(v.w, hi.w);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// src (N, C, HW) NCHW -> S (N, HW, ctot): 32(c) x 32(p) tiles through shared memory
__global__ void __launch_bounds__(256)
gru_stage_kernel(const float* __restrict__ src, int C, long long hw, float* __restrict__ S, int ctot, int off_hi0, int off_lo,
                 int off_hi1) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const long long p0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    const long long p = p0 + tx;
    tile[ty + 8 * i][tx] = (c < C && p < hw) ? __ldg(src + (n * C + c) * hw + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p = p0 + ty + 8 * i;
    const int c = c0 + tx;
    if (c < C && p < hw) {
      const float v = tile[tx][ty + 8 * i];
      const float hi = rn_tf32_g(v);
      float* row = S + (n * hw + p) * ctot;
      row[off_hi0 + c] = hi;
      row[off_lo + c] = __fsub_rn(v, hi);
      row[off_hi1 + c] = hi;
    }
  }
}

// channels-last source (N, HW, C): no transpose, one thread = 4 consecutive channels of one pixel
__global__ void __launch_bounds__(256)
gru_stage_cl_kernel(const float4* __restrict__ src, int c4n, long long total4, float* __restrict__ S, int ctot, int off_hi0,
                    int off_lo, int off_hi1) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / c4n;
    const int c4 = static_cast<int>(i - p * c4n);
    float4 hi, lo;
    split4(__ldg(src + i), hi, lo);
    float* row = S + p * ctot + 4 * c4;
    *reinterpret_cast<float4*>(row + off_hi0) = hi;
    *reinterpret_cast<float4*>(row + off_lo) = lo;
    *reinterpret_cast<float4*>(row + off_hi1) = hi;
  }
}

// one thread = 4 consecutive channels of one pixel; everything channels-last
__global__ void __launch_bounds__(256)
gru_gate_r_kernel(const float4* __restrict__ zr_pre, const float4* __restrict__ bias_zr, const float4* __restrict__ h, int ch4,
                  long long total4, float4* __restrict__ z_out, float* __restrict__ S, int ctot) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / ch4;
    const int c4 = static_cast<int>(i - p * ch4);
    const float4 zp = __ldg(zr_pre + p * 2 * ch4 + c4), rp = __ldg(zr_pre + p * 2 * ch4 + ch4 + c4);
    const float4 bz = __ldg(bias_zr + c4), br = __ldg(bias_zr + ch4 + c4);
    const float4 hv = __ldg(h + i);
    float4 z, rh;
    z.x = sigmoidf_(zp.x + bz.x); z.y = sigmoidf_(zp.y + bz.y); z.z = sigmoidf_(zp.z + bz.z); z.w = sigmoidf_(zp.w + bz.w);
    rh.x = sigmoidf_(rp.x + br.x) * hv.x; rh.y = sigmoidf_(rp.y + br.y) * hv.y;
    rh.z = sigmoidf_(rp.z + br.z) * hv.z; rh.w = sigmoidf_(rp.w + br.w) * hv.w;
    z_out[i] = z;
    float4 hi, lo;
    split4(rh, hi, lo);
    float4* row = reinterpret_cast<float4*>(S + p * ctot);
    row[c4] = hi;
    row[ch4 + c4] = lo;
    row[2 * ch4 + c4] = hi;
  }
}

__global__ void __launch_bounds__(256)
gru_gate_h_kernel(const float4* __restrict__ q_pre, const float4* __restrict__ bias_q, const float4* __restrict__ z, int ch4,
                  long long total4, float4* __restrict__ h, float* __restrict__ S, int ctot) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / ch4;
    const int c4 = static_cast<int>(i - p * ch4);
    const float4 qp = __ldg(q_pre + i), bq = __ldg(bias_q + c4), zv = __ldg(z + i), hv = h[i];
    float4 hn;
    hn.x = (1.0f - zv.x) * hv.x + zv.x * tanhf(qp.x + bq.x);
    hn.y = (1.0f - zv.y) * hv.y + zv.y * tanhf(qp.y + bq.y);
    hn.z = (1.0f - zv.z) * hv.z + zv.z * tanhf(qp.z + bq.z);
    hn.w = (1.0f - zv.w) * hv.w + zv.w * tanhf(qp.w + bq.w);
    h[i] = hn;
    float4 hi, lo;
    split4(hn, hi, lo);
    float4* row = reinterpret_cast<float4*>(S + p * ctot);
    row[c4] = hi;
    row[ch4 + c4] = lo;
    row[2 * ch4 + c4] = hi;
  }
}

static unsigned grid_for(long long items) {
  const long long want = (items + 255) / 256, cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<unsigned>(want < cap ? want : cap);
}

}  // namespace nnd

extern "C" {

nnd_status nnd_gru_stage(const float* src, int src_channels_last, int N, int C, long long hw, float* S, int ctot, int off_hi0,
                         int off_lo, int off_hi1, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(src && S, "gru_stage: null pointer");
  NND_REQUIRE(N > 0 && C > 0 && hw > 0 && ctot > 0, "gru_stage: N, C, H*W, ctot must be positive");
  NND_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, "gru_stage: N or C exceeds the grid limit");
  NND_REQUIRE(off_hi0 >= 0 && off_lo >= 0 && off_hi1 >= 0 && off_hi0 + C <= ctot && off_lo + C <= ctot && off_hi1 + C <= ctot,
              "gru_stage: channel offsets outside the staging row");
  if (src_channels_last) {
    NND_REQUIRE(C % 4 == 0 && ctot % 4 == 0 && off_hi0 % 4 == 0 && off_lo % 4 == 0 && off_hi1 % 4 == 0 && aligned16(src) &&
                    aligned16(S),
                "gru_stage: the channels-last source path needs channel counts / offsets in quads and 16-byte alignment");
    const long long total4 = static_cast<long long>(N) * hw * (C / 4);
    gru_stage_cl_kernel<<<grid_for(total4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(src), C / 4, total4, S, ctot, off_hi0, off_lo, off_hi1);
    return check_launch("gru_stage_cl_kernel");
  }
  dim3 grid(static_cast<unsigned>((hw + 31) / 32), (C + 31) / 32, N);
  gru_stage_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, C, hw, S, ctot, off_hi0, off_lo, off_hi1);
  return check_launch("gru_stage_kernel");
}

nnd_status nnd_gru_gate_r(const float* zr_pre, const float* bias_zr, const float* h, long long pixels, int ch, float* z,
                          float* S, int ctot, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(zr_pre && bias_zr && h && z && S, "gru_gate_r: null pointer");
  NND_REQUIRE(pixels > 0 && ch > 0 && ch % 4 == 0 && ctot % 4 == 0 && 3 * ch <= ctot,
              "gru_gate_r: needs ch %% 4 == 0, ctot %% 4 == 0 and 3*ch <= ctot");
  NND_REQUIRE(aligned16(zr_pre) && aligned16(bias_zr) && aligned16(h) && aligned16(z) && aligned16(S),
              "gru_gate_r: tensors must be 16-byte aligned");
  const long long total4 = pixels * (ch / 4);
  gru_gate_r_kernel<<<grid_for(total4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(zr_pre), reinterpret_cast<const float4*>(bias_zr), reinterpret_cast<const float4*>(h),
      ch / 4, total4, reinterpret_cast<float4*>(z), S, ctot);
  return check_launch("gru_gate_r_kernel");
}

nnd_status nnd_gru_gate_h(const float* q_pre, const float* bias_q, const float* z, long long pixels, int ch, float* h,
                          float* S, int ctot, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(q_pre && bias_q && z && h && S, "gru_gate_h: null pointer");
  NND_REQUIRE(pixels > 0 && ch > 0 && ch % 4 == 0 && ctot % 4 == 0 && 3 * ch <= ctot,
              "gru_gate_h: needs ch %% 4 == 0, ctot %% 4 == 0 and 3*ch <= ctot");
  NND_REQUIRE(aligned16(q_pre) && aligned16(bias_q) && aligned16(z) && aligned16(h) && aligned16(S),
              "gru_gate_h: tensors must be 16-byte aligned");
  const long long total4 = pixels * (ch / 4);
  gru_gate_h_kernel<<<grid_for(total4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(q_pre), reinterpret_cast<const float4*>(bias_q), reinterpret_cast<const float4*>(z), ch / 4,
      total4, reinterpret_cast<float4*>(h), S, ctot);
  return check_launch("gru_gate_h_kernel");
}

}  // extern "C"
