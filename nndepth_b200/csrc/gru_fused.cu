// Elementwise glue of the separable ConvGRU around its (cuDNN, TF32 tensor-core) convolutions, fused and
// channels-last (sm_100a).  Reference: nndepth/blocks/gru.py:5-37
//     hx = cat[h, x];  z = sigmoid(convz(hx));  r = sigmoid(convr(hx));
//     q = tanh(convq(cat[r*h, x]));  h = (1 - z) * h + z * q            (twice: 1x5 then 5x1 kernels)
//
// Precision: plain TF32 convolutions in this recurrence move the final disparity 0.013 px from the fp32
// reference (bar 0.01 px).  The culprit is the rounding of the WEIGHTS -- the same perturbation is applied in
// all 32 iterations and accumulates coherently -- not of the activations, whose rounding errors are fresh every
// iteration and average out (tools/exp_epe_2term.py: activations rounded + weights exact 0.0031 px; weights
// rounded + activations exact 0.0113 px).  So the convolutions run as
//     conv(RN_tf32(x), w_hi) + conv(RN_tf32(x), w_lo),   w_hi = RN_tf32(w), w_lo = w - w_hi,
// i.e. ONE TF32 convolution with the output channels doubled ([w_hi ; w_lo]) whose two halves are added by the
// gate kernels below: fp32-exact weights on the tensor cores at 2x (not 4x fp32-CUDA-core, not 3x "3xTF32")
// the cost of a TF32 convolution.
//
// Unfused, every half-step makes ~15 passes over its tensors (cat, cuDNN's NCHW->NHWC conversion, bias add,
// sigmoid, chunk, mul, cat, conversion, bias add, tanh, rsub, mul, mul, add).  Here one channels-last staging
// buffer   S[n][p][ RN(h) (ch) | RN(x) (cx) | RN(h) | RN(x) ]   is the NHWC input of all four convolutions of an
// iteration (no layout conversions: cuDNN's tensor-core kernels are NHWC-native): the x part is written once per iteration
// (the context half of x once per forward), the h part is overwritten in place -- by r*h for the q
// convolution, by the new h for the next half-step.
//
//   nnd_gru_stage    NCHW or NHWC tensor -> RN_tf32 at a channel offset of both halves of S   (transpose + round)
//   nnd_gru_gate_r   zr_pre (NHWC, [z | r]) + bias, h -> z ; S.h <- RN(r * h)
//   nnd_gru_gate_h   q_pre (NHWC) + bias, z, h -> h' = (1-z) h + z tanh(q) ; S.h <- RN(h') ; h <- h'
#include "common.cuh"

namespace nnd {

__device__ __forceinline__ float rn_tf32_g(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return __uint_as_float(y);
}

__device__ __forceinline__ float4 rn4(const float4 v) {
  return make_float4(rn_tf32_g(v.x), rn_tf32_g(v.y), rn_tf32_g(v.z), rn_tf32_g(v.w));
}

__device__ __forceinline__ float4 add4(const float4 a, const float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// src (N, C, HW) NCHW -> S (N, HW, ctot) at channel offset `off`: 32(c) x 32(p) tiles through shared memory
__global__ void __launch_bounds__(256)
gru_stage_kernel(const float* __restrict__ src, int C, long long hw, float* __restrict__ S, int ctot, int off) {
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
      const float v = rn_tf32_g(tile[tx][ty + 8 * i]);
      float* row = S + (n * hw + p) * ctot + off + c;
      row[0] = v;
      row[ctot / 2] = v;
    }
  }
}

// channels-last source (N, HW, C): no transpose, one thread = 4 consecutive channels of one pixel
__global__ void __launch_bounds__(256)
gru_stage_cl_kernel(const float4* __restrict__ src, int c4n, long long total4, float* __restrict__ S, int ctot, int off) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / c4n;
    const int c4 = static_cast<int>(i - p * c4n);
    const float4 v = rn4(__ldg(src + i));
    float* row = S + p * ctot + off + 4 * c4;
    *reinterpret_cast<float4*>(row) = v;
    *reinterpret_cast<float4*>(row + ctot / 2) = v;
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
    const float4* row = zr_pre + p * 2 * ch4;   // [z | r]
    const float4 zp = add4(__ldg(row + c4), __ldg(bias_zr + c4));
    const float4 rp = add4(__ldg(row + ch4 + c4), __ldg(bias_zr + ch4 + c4));
    const float4 hv = __ldg(h + i);
    z_out[i] = make_float4(sigmoidf_(zp.x), sigmoidf_(zp.y), sigmoidf_(zp.z), sigmoidf_(zp.w));
    const float4 rh = make_float4(sigmoidf_(rp.x) * hv.x, sigmoidf_(rp.y) * hv.y, sigmoidf_(rp.z) * hv.z, sigmoidf_(rp.w) * hv.w);
    const float4 rq = rn4(rh);
    *reinterpret_cast<float4*>(S + p * ctot + 4 * c4) = rq;
    *reinterpret_cast<float4*>(S + p * ctot + ctot / 2 + 4 * c4) = rq;
  }
}

__global__ void __launch_bounds__(256)
gru_gate_h_kernel(const float4* __restrict__ q_pre, const float4* __restrict__ bias_q, const float4* __restrict__ z, int ch4,
                  long long total4, float4* __restrict__ h, float* __restrict__ S, int ctot) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / ch4;
    const int c4 = static_cast<int>(i - p * ch4);
    const float4 qp = add4(__ldg(q_pre + i), __ldg(bias_q + c4));
    const float4 zv = __ldg(z + i), hv = h[i];
    float4 hn;
    hn.x = (1.0f - zv.x) * hv.x + zv.x * tanhf(qp.x);
    hn.y = (1.0f - zv.y) * hv.y + zv.y * tanhf(qp.y);
    hn.z = (1.0f - zv.z) * hv.z + zv.z * tanhf(qp.z);
    hn.w = (1.0f - zv.w) * hv.w + zv.w * tanhf(qp.w);
    h[i] = hn;
    const float4 hq = rn4(hn);
    *reinterpret_cast<float4*>(S + p * ctot + 4 * c4) = hq;
    *reinterpret_cast<float4*>(S + p * ctot + ctot / 2 + 4 * c4) = hq;
  }
}

// ---- fp16 staging variant ------------------------------------------------------------------------------------
// fp16 has TF32's 10-bit mantissa, so RN_fp16(x) == RN_tf32(x) for every |x| in fp16's normal range (the GRU's
// activations: h in [-1, 1], x = ReLU features of a few units) and kind::f16 tensor-core products run at twice
// the TF32 rate.  The staging buffer, the split weights [RN16(w) ; RN16(w - RN16(w))] and the pre-activations
// cuDNN returns are fp16; accumulation is fp32 inside the convolution; h, z and all gate arithmetic stay fp32.
// Conversions saturate (cvt.rn.satfinite) instead of producing inf.  Measured (tools/exp_epe_fp16.py): final
// disparity 0.00312 px from the reference vs 0.00311 px for the TF32 split.
__device__ __forceinline__ void store_both_h4(uint16_t* S, long long p, int ctot, int c, const uint2 v) {
  *reinterpret_cast<uint2*>(S + p * ctot + c) = v;
  *reinterpret_cast<uint2*>(S + p * ctot + ctot / 2 + c) = v;
}

// fp32 NCHW source -> fp16 staging rows (transpose through shared memory, as gru_stage_kernel)
__global__ void __launch_bounds__(256)
gru_stage_f16_kernel(const float* __restrict__ src, int C, long long hw, uint16_t* __restrict__ S, int ctot, int off) {
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
      const uint16_t v = static_cast<uint16_t>(pack_h2(tile[tx][ty + 8 * i], 0.f) & 0xffffu);
      uint16_t* row = S + (n * hw + p) * ctot + off + c;
      row[0] = v;
      row[ctot / 2] = v;
    }
  }
}

// channels-last source, fp32 (SRC_HALF = 0) or fp16 (1): one thread = 4 consecutive channels of one pixel
template <int SRC_HALF>
__global__ void __launch_bounds__(256)
gru_stage_cl_f16_kernel(const void* __restrict__ src, int c4n, long long total4, uint16_t* __restrict__ S, int ctot, int off) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / c4n;
    const int c4 = static_cast<int>(i - p * c4n);
    const uint2 v = SRC_HALF ? __ldg(reinterpret_cast<const uint2*>(src) + i)
                             : pack_h4(__ldg(reinterpret_cast<const float4*>(src) + i));
    store_both_h4(S, p, ctot, off + 4 * c4, v);
  }
}

__global__ void __launch_bounds__(256)
gru_gate_r_f16_kernel(const uint2* __restrict__ zr_pre, const float4* __restrict__ bias_zr, const float4* __restrict__ h,
                      int ch4, long long total4, float4* __restrict__ z_out, uint16_t* __restrict__ S, int ctot) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / ch4;
    const int c4 = static_cast<int>(i - p * ch4);
    const uint2* row = zr_pre + p * 2 * ch4;   // [z | r]
    const float4 zp = add4(unpack_h4(__ldg(row + c4)), __ldg(bias_zr + c4));
    const float4 rp = add4(unpack_h4(__ldg(row + ch4 + c4)), __ldg(bias_zr + ch4 + c4));
    const float4 hv = __ldg(h + i);
    z_out[i] = make_float4(sigmoidf_(zp.x), sigmoidf_(zp.y), sigmoidf_(zp.z), sigmoidf_(zp.w));
    const float4 rh = make_float4(sigmoidf_(rp.x) * hv.x, sigmoidf_(rp.y) * hv.y, sigmoidf_(rp.z) * hv.z, sigmoidf_(rp.w) * hv.w);
    store_both_h4(S, p, ctot, 4 * c4, pack_h4(rh));
  }
}

__global__ void __launch_bounds__(256)
gru_gate_h_f16_kernel(const uint2* __restrict__ q_pre, const float4* __restrict__ bias_q, const float4* __restrict__ z, int ch4,
                      long long total4, float4* __restrict__ h, uint16_t* __restrict__ S, int ctot, uint2* __restrict__ h16) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / ch4;
    const int c4 = static_cast<int>(i - p * ch4);
    const float4 qp = add4(unpack_h4(__ldg(q_pre + i)), __ldg(bias_q + c4));
    const float4 zv = __ldg(z + i), hv = h[i];
    float4 hn;
    hn.x = (1.0f - zv.x) * hv.x + zv.x * tanhf(qp.x);
    hn.y = (1.0f - zv.y) * hv.y + zv.y * tanhf(qp.y);
    hn.z = (1.0f - zv.z) * hv.z + zv.z * tanhf(qp.z);
    hn.w = (1.0f - zv.w) * hv.w + zv.w * tanhf(qp.w);
    h[i] = hn;
    const uint2 hq = pack_h4(hn);
    store_both_h4(S, p, ctot, 4 * c4, hq);
    if (h16) h16[i] = hq;
  }
}

static unsigned grid_for(long long items) {
  const long long want = (items + 255) / 256, cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<unsigned>(want < cap ? want : cap);
}

}  // namespace nnd

extern "C" {

nnd_status nnd_gru_stage(const float* src, int src_channels_last, int N, int C, long long hw, float* S, int ctot, int off,
                         nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(src && S, "gru_stage: null pointer");
  NND_REQUIRE(N > 0 && C > 0 && hw > 0 && ctot > 0, "gru_stage: N, C, H*W, ctot must be positive");
  NND_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, "gru_stage: N or C exceeds the grid limit");
  NND_REQUIRE(ctot % 2 == 0 && off >= 0 && off + C <= ctot / 2, "gru_stage: channel offset outside the staging half-row");
  if (src_channels_last) {
    NND_REQUIRE(C % 4 == 0 && ctot % 8 == 0 && off % 4 == 0 && aligned16(src) && aligned16(S),
                "gru_stage: the channels-last source path needs channel counts / offsets in quads and 16-byte alignment");
    const long long total4 = static_cast<long long>(N) * hw * (C / 4);
    gru_stage_cl_kernel<<<grid_for(total4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(src), C / 4, total4, S, ctot, off);
    return check_launch("gru_stage_cl_kernel");
  }
  dim3 grid(static_cast<unsigned>((hw + 31) / 32), (C + 31) / 32, N);
  gru_stage_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, C, hw, S, ctot, off);
  return check_launch("gru_stage_kernel");
}

nnd_status nnd_gru_gate_r(const float* zr_pre, const float* bias_zr, const float* h, long long pixels, int ch, float* z,
                          float* S, int ctot, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(zr_pre && bias_zr && h && z && S, "gru_gate_r: null pointer");
  NND_REQUIRE(pixels > 0 && ch > 0 && ch % 4 == 0 && ctot % 8 == 0 && 2 * ch <= ctot,
              "gru_gate_r: needs ch %% 4 == 0, ctot %% 8 == 0 and 2*ch <= ctot");
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
  NND_REQUIRE(pixels > 0 && ch > 0 && ch % 4 == 0 && ctot % 8 == 0 && 2 * ch <= ctot,
              "gru_gate_h: needs ch %% 4 == 0, ctot %% 8 == 0 and 2*ch <= ctot");
  NND_REQUIRE(aligned16(q_pre) && aligned16(bias_q) && aligned16(z) && aligned16(h) && aligned16(S),
              "gru_gate_h: tensors must be 16-byte aligned");
  const long long total4 = pixels * (ch / 4);
  gru_gate_h_kernel<<<grid_for(total4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(q_pre), reinterpret_cast<const float4*>(bias_q), reinterpret_cast<const float4*>(z), ch / 4,
      total4, reinterpret_cast<float4*>(h), S, ctot);
  return check_launch("gru_gate_h_kernel");
}

nnd_status nnd_gru_stage_f16(const void* src, int src_kind, int N, int C, long long hw, void* S16, int ctot, int off,
                             nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(src && S16, "gru_stage_f16: null pointer");
  NND_REQUIRE(src_kind >= 0 && src_kind <= 2, "gru_stage_f16: src_kind %d is not 0 (fp32 NCHW), 1 (fp32 NHWC) or 2 (fp16 NHWC)",
              src_kind);
  NND_REQUIRE(N > 0 && C > 0 && hw > 0 && ctot > 0, "gru_stage_f16: N, C, H*W, ctot must be positive");
  NND_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, "gru_stage_f16: N or C exceeds the grid limit");
  NND_REQUIRE(ctot % 2 == 0 && off >= 0 && off + C <= ctot / 2, "gru_stage_f16: channel offset outside the staging half-row");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint16_t* S = reinterpret_cast<uint16_t*>(S16);
  if (src_kind != 0) {
    NND_REQUIRE(C % 4 == 0 && ctot % 8 == 0 && off % 4 == 0 && aligned16(src) && aligned16(S16),
                "gru_stage_f16: the channels-last source path needs channel counts / offsets in quads and 16-byte alignment");
    const long long total4 = static_cast<long long>(N) * hw * (C / 4);
    if (src_kind == 1) {
      gru_stage_cl_f16_kernel<0><<<grid_for(total4), 256, 0, st>>>(src, C / 4, total4, S, ctot, off);
    } else {
      gru_stage_cl_f16_kernel<1><<<grid_for(total4), 256, 0, st>>>(src, C / 4, total4, S, ctot, off);
    }
    return check_launch("gru_stage_cl_f16_kernel");
  }
  dim3 grid(static_cast<unsigned>((hw + 31) / 32), (C + 31) / 32, N);
  gru_stage_f16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(src), C, hw, S, ctot, off);
  return check_launch("gru_stage_f16_kernel");
}

nnd_status nnd_gru_gate_r_f16(const void* zr_pre16, const float* bias_zr, const float* h, long long pixels, int ch, float* z,
                              void* S16, int ctot, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(zr_pre16 && bias_zr && h && z && S16, "gru_gate_r_f16: null pointer");
  NND_REQUIRE(pixels > 0 && ch > 0 && ch % 4 == 0 && ctot % 8 == 0 && 2 * ch <= ctot,
              "gru_gate_r_f16: needs ch %% 4 == 0, ctot %% 8 == 0 and 2*ch <= ctot");
  NND_REQUIRE(aligned16(zr_pre16) && aligned16(bias_zr) && aligned16(h) && aligned16(z) && aligned16(S16),
              "gru_gate_r_f16: tensors must be 16-byte aligned");
  const long long total4 = pixels * (ch / 4);
  gru_gate_r_f16_kernel<<<grid_for(total4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint2*>(zr_pre16), reinterpret_cast<const float4*>(bias_zr), reinterpret_cast<const float4*>(h),
      ch / 4, total4, reinterpret_cast<float4*>(z), reinterpret_cast<uint16_t*>(S16), ctot);
  return check_launch("gru_gate_r_f16_kernel");
}

nnd_status nnd_gru_gate_h_f16(const void* q_pre16, const float* bias_q, const float* z, long long pixels, int ch, float* h,
                              void* S16, int ctot, void* h16, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(q_pre16 && bias_q && z && h && S16, "gru_gate_h_f16: null pointer");
  NND_REQUIRE(pixels > 0 && ch > 0 && ch % 4 == 0 && ctot % 8 == 0 && 2 * ch <= ctot,
              "gru_gate_h_f16: needs ch %% 4 == 0, ctot %% 8 == 0 and 2*ch <= ctot");
  NND_REQUIRE(aligned16(q_pre16) && aligned16(bias_q) && aligned16(z) && aligned16(h) && aligned16(S16) && aligned16(h16),
              "gru_gate_h_f16: tensors must be 16-byte aligned");
  const long long total4 = pixels * (ch / 4);
  gru_gate_h_f16_kernel<<<grid_for(total4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint2*>(q_pre16), reinterpret_cast<const float4*>(bias_q), reinterpret_cast<const float4*>(z), ch / 4,
      total4, reinterpret_cast<float4*>(h), reinterpret_cast<uint16_t*>(S16), ctot, reinterpret_cast<uint2*>(h16));
  return check_launch("gru_gate_h_f16_kernel");
}

}  // extern "C"
