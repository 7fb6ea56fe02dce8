// The two single-flow-channel convolutions of the RAFT-Stereo update block (sm_100a).  With one flow channel
// (stereo) neither is GEMM-shaped: cuDNN pads the channel to its tile width, runs a tensor-core kernel that is
// >97 % padding and converts layouts around it (3 + 49 + 6 us and 42 us per iteration at KITTI shapes).
//
//   nnd_flow_conv7x7_relu   relu(convf1(flow)):  (N,1,H,W) -> channels-last (N,H,W,Cout), 7x7, padding 3
//                           reference blocks/update_block.py:53,60
//   nnd_flow_head_tail      FlowHead.conv2 (3x3, C -> 1) on a channels-last map, fused with the coordinate update
//                           of the refinement loop: delta = conv2(x); coords += delta; flow = coords - org
//                           reference blocks/update_block.py:23,36 and raft_stereo/model.py:132-134
//
// Both accumulate in fp32 FFMA (cuDNN's TF32 kernels would round the operands; the reference is fp32).
#include "common.cuh"

namespace nnd {

constexpr int F7_PX = 8;  // consecutive x positions per thread (share every weight load)

// block = (Cout/4 lanes-of-4-channels, rows of 8-pixel strips); weight_t = the filter bank as [tap][Cout]
template <int OUT_F16>
__global__ void __launch_bounds__(256)
flow_conv7x7_relu_kernel(const float* __restrict__ flow, const float* __restrict__ weight_t, const float* __restrict__ bias,
                         int H, int W, int cout, long long n_strips, int strips_per_row, void* __restrict__ out) {
  extern __shared__ float4 w_sm[];  // [49][cout / 4]
  const int c4n = cout >> 2;
  for (int i = threadIdx.x; i < 49 * c4n; i += blockDim.x) w_sm[i] = __ldg(reinterpret_cast<const float4*>(weight_t) + i);
  __syncthreads();
  const int c4 = threadIdx.x % c4n;
  const int sub = threadIdx.x / c4n, subs = blockDim.x / c4n;
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + c4);
  for (long long s = static_cast<long long>(blockIdx.x) * subs + sub; s < n_strips;
       s += static_cast<long long>(gridDim.x) * subs) {
    const long long row = s / strips_per_row;  // n * H + y
    const int x0 = static_cast<int>(s - row * strips_per_row) * F7_PX;
    const int y = static_cast<int>(row % H);
    const float* img = flow + (row - y) * W;   // image n
    float2 acc[F7_PX][2];   // channels (x, y) and (z, w) of each pixel: packed FFMA2 accumulators
#pragma unroll
    for (int p = 0; p < F7_PX; ++p) {
      acc[p][0] = make_float2(b4.x, b4.y);
      acc[p][1] = make_float2(b4.z, b4.w);
    }
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
      const int yy = y + ky - 3;
      if (yy < 0 || yy >= H) continue;
      float f[F7_PX + 6];
      const float* line = img + static_cast<long long>(yy) * W + x0 - 3;
      if (x0 >= 3 && x0 + F7_PX + 3 <= W) {   // interior strip (warp-uniform): immediate offsets, no predicates
#pragma unroll
        for (int i = 0; i < F7_PX + 6; ++i) f[i] = __ldg(line + i);
      } else {
#pragma unroll
        for (int i = 0; i < F7_PX + 6; ++i) {
          const int xx = x0 + i - 3;
          f[i] = (xx >= 0 && xx < W) ? __ldg(line + i) : 0.f;
        }
      }
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float4 w4 = w_sm[(ky * 7 + kx) * c4n + c4];
        const float2 wlo = make_float2(w4.x, w4.y), whi = make_float2(w4.z, w4.w);
#pragma unroll
        for (int p = 0; p < F7_PX; ++p) {
          const float2 ff = make_float2(f[p + kx], f[p + kx]);
          ffma2(acc[p][0], wlo, ff);
          ffma2(acc[p][1], whi, ff);
        }
      }
    }
    const long long q0 = (row * W + x0) * c4n + c4;   // index in units of 4 channels
#pragma unroll
    for (int p = 0; p < F7_PX; ++p) {
      if (x0 + p < W) {
        const float4 v = make_float4(fmaxf(acc[p][0].x, 0.f), fmaxf(acc[p][0].y, 0.f), fmaxf(acc[p][1].x, 0.f),
                                     fmaxf(acc[p][1].y, 0.f));
        if (OUT_F16) {
          reinterpret_cast<uint2*>(out)[q0 + static_cast<long long>(p) * c4n] = pack_h4(v);
        } else {
          reinterpret_cast<float4*>(out)[q0 + static_cast<long long>(p) * c4n] = v;
        }
      }
    }
  }
}

constexpr int FH_PX = 8;  // output pixels per warp (a strip along x)

template <int F16>
struct QuadOf {
  using type = float4;
};
template <>
struct QuadOf<1> {
  using type = uint2;
};
__device__ __forceinline__ float4 quad_to_float4(const float4 v) { return v; }
__device__ __forceinline__ float4 quad_to_float4(const uint2 v) { return unpack_h4(v); }

// x (N,H,W,C) channels-last with C = 32 * CPL: lane owns channels [lane*CPL, lane*CPL + CPL).  weight (1, C, 3, 3).
// One warp per strip of FH_PX output pixels: the 3 x (FH_PX + 2) input pixels are read once each.
template <int CPL, int X_F16>
__global__ void __launch_bounds__(256)
flow_head_tail_kernel(const void* __restrict__ x_, const float* __restrict__ weight, const float* __restrict__ bias,
                      int H, int W, long long n_strips, int strips_per_row, float* __restrict__ delta,
                      const float* __restrict__ coords_in, const float* __restrict__ org, float* __restrict__ coords_out,
                      float* __restrict__ flow_out) {
  using RawQuad = typename QuadOf<X_F16>::type;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  constexpr int C = 32 * CPL;
  // weights: read coalesced once per block, handed to the lanes through shared memory laid out [tap][j][lane]
  // (a lane's 9 * CPL values straight from global memory are 9 * CPL scattered 4-byte reads per warp: 283 MB of
  // L2 traffic per launch at KITTI shapes for 16 MB of activations)
  __shared__ float w_s[9 * CPL * 32];
  for (int idx = threadIdx.x; idx < 9 * C; idx += blockDim.x) {
    const int c = idx / 9, t = idx - 9 * c;
    w_s[(t * CPL + (c % CPL)) * 32 + c / CPL] = __ldg(weight + idx);
  }
  __syncthreads();
  float w[9][CPL];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < CPL; ++j) w[t][j] = w_s[(t * CPL + j) * 32 + lane];
  const float b = bias ? __ldg(bias) : 0.f;
  for (long long s = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); s < n_strips;
       s += static_cast<long long>(gridDim.x) * warps_per_block) {
    const long long row = s / strips_per_row;  // n * H + y
    const int x0 = static_cast<int>(s - row * strips_per_row) * FH_PX;
    const int y = static_cast<int>(row % H);
    float acc[FH_PX];
#pragma unroll
    for (int p = 0; p < FH_PX; ++p) acc[p] = 0.f;
    // loads are predicated, not branched around, and batched: with 4 channels per lane all 3 x (FH_PX + 2) of a strip
    // are issued back to back before the first FMA (one DRAM latency per strip); wider lanes batch one row at a time
    constexpr int KYB = CPL == 4 ? 3 : 1;
#pragma unroll
    for (int ky0 = 0; ky0 < 3; ky0 += KYB) {
      RawQuad v4[KYB][FH_PX + 2][CPL / 4];   // raw bits: an fp16 quad stays 8 bytes until the FMAs need it
#pragma unroll
      for (int kb = 0; kb < KYB; ++kb) {
        const int yy = y + ky0 + kb - 1;
        const bool row_ok = yy >= 0 && yy < H;
        const long long line = ((row - y + (row_ok ? yy : y)) * W) * C + lane * CPL;   // element index of this lane's channels
#pragma unroll
        for (int i = 0; i < FH_PX + 2; ++i) {
          const int xx = x0 + i - 1;
          const bool ok = row_ok && xx >= 0 && xx < W;
#pragma unroll
          for (int q = 0; q < CPL / 4; ++q) {
            RawQuad a;
            memset(&a, 0, sizeof(a));
            if (ok) a = __ldg(reinterpret_cast<const RawQuad*>(x_) + (line + static_cast<long long>(xx) * C) / 4 + q);
            v4[kb][i][q] = a;
          }
        }
      }
#pragma unroll
      for (int kb = 0; kb < KYB; ++kb) {
        const int ky = ky0 + kb;
#pragma unroll
        for (int i = 0; i < FH_PX + 2; ++i) {
          float v[CPL];
#pragma unroll
          for (int q = 0; q < CPL / 4; ++q) {
            const float4 f = quad_to_float4(v4[kb][i][q]);
            v[4 * q] = f.x;
            v[4 * q + 1] = f.y;
            v[4 * q + 2] = f.z;
            v[4 * q + 3] = f.w;
          }
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int p = i - kx;  // output pixel that sees input i through tap kx
            if (p < 0 || p >= FH_PX) continue;
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[p] = fmaf(v[j], w[ky * 3 + kx][j], acc[p]);
          }
        }
      }
    }
    // 8 butterfly reductions; lane p keeps pixel p
    float mine = 0.f;
#pragma unroll
    for (int p = 0; p < FH_PX; ++p) {
      float v = acc[p];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == p) mine = v;
    }
    if (lane < FH_PX && x0 + lane < W) {
      const long long idx = row * W + x0 + lane;
      const float d = mine + b;
      if (delta) delta[idx] = d;
      if (coords_in) {
        const float c1 = coords_in[idx] + d;
        coords_out[idx] = c1;
        if (flow_out) flow_out[idx] = c1 - org[idx];
      }
    }
  }
}

// channels-last concatenation with conversion to fp16: out[p] = [a[p] (Ca) | b[p] (Cb)], each source fp32 or fp16.
// One thread = 4 channels of one pixel.
__global__ void __launch_bounds__(256)
nhwc_cat_f16_kernel(const void* __restrict__ a, int a_f16, int ca4, const void* __restrict__ b, int b_f16, int cb4,
                    long long total4, uint2* __restrict__ out) {
  const int ct4 = ca4 + cb4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / ct4;
    const int c4 = static_cast<int>(i - p * ct4);
    const bool from_a = c4 < ca4;
    const void* src = from_a ? a : b;
    const long long idx = from_a ? p * ca4 + c4 : p * cb4 + (c4 - ca4);
    const int f16 = from_a ? a_f16 : b_f16;
    out[i] = f16 ? __ldg(reinterpret_cast<const uint2*>(src) + idx) : pack_h4(__ldg(reinterpret_cast<const float4*>(src) + idx));
  }
}

}  // namespace nnd

extern "C" {

nnd_status nnd_nhwc_cat_f16(const void* a, int a_f16, int c_a, const void* b, int b_f16, int c_b, long long pixels, void* out,
                            nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(a && b && out, "nhwc_cat_f16: null pointer");
  NND_REQUIRE(pixels > 0 && c_a > 0 && c_b > 0 && c_a % 4 == 0 && c_b % 4 == 0,
              "nhwc_cat_f16: channel counts must be positive multiples of 4");
  NND_REQUIRE(aligned16(a) && aligned16(b) && aligned16(out), "nhwc_cat_f16: tensors must be 16-byte aligned");
  const long long total4 = pixels * ((c_a + c_b) / 4);
  const long long want = (total4 + 255) / 256, cap = static_cast<long long>(sm_count()) * 16;
  nhwc_cat_f16_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      a, a_f16, c_a / 4, b, b_f16, c_b / 4, total4, reinterpret_cast<uint2*>(out));
  return check_launch("nhwc_cat_f16_kernel");
}

nnd_status nnd_flow_conv7x7_relu(const float* flow, const float* weight_t, const float* bias, int N, int H, int W,
                                 int c_out, void* out, int out_f16, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(flow && weight_t && bias && out, "flow_conv7x7_relu: null pointer");
  NND_REQUIRE(N > 0 && H > 0 && W > 0, "flow_conv7x7_relu: N, H, W must be positive");
  NND_REQUIRE(c_out > 0 && c_out % 4 == 0 && c_out <= 1024 && 256 % (c_out / 4) == 0,
              "flow_conv7x7_relu: c_out = %d must be a multiple of 4 with c_out/4 dividing 256", c_out);
  NND_REQUIRE(aligned16(out) && aligned16(bias) && aligned16(weight_t),
              "flow_conv7x7_relu: out, bias and weight_t must be 16-byte aligned");
  const int strips_per_row = (W + F7_PX - 1) / F7_PX;
  const long long n_strips = static_cast<long long>(N) * H * strips_per_row;
  const int subs = 256 / (c_out / 4);
  const long long want = (n_strips + subs - 1) / subs;
  const long long cap = static_cast<long long>(sm_count()) * 2;  // persistent: the 49 x c_out weights are staged per block
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  const size_t smem = static_cast<size_t>(49) * c_out * sizeof(float);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_f16) {
    flow_conv7x7_relu_kernel<1><<<grid, 256, smem, st>>>(flow, weight_t, bias, H, W, c_out, n_strips, strips_per_row, out);
  } else {
    flow_conv7x7_relu_kernel<0><<<grid, 256, smem, st>>>(flow, weight_t, bias, H, W, c_out, n_strips, strips_per_row, out);
  }
  return check_launch("flow_conv7x7_relu_kernel");
}

nnd_status nnd_flow_head_tail(const void* x, int x_f16, const float* weight, const float* bias, int N, int C, int H, int W,
                              float* delta, const float* coords_in, const float* org, float* coords_out,
                              float* flow_out, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(x && weight, "flow_head_tail: null pointer");
  NND_REQUIRE(N > 0 && H > 0 && W > 0, "flow_head_tail: N, H, W must be positive");
  NND_REQUIRE(C == 128 || C == 256 || C == 512, "flow_head_tail: C = %d (supported: 128, 256, 512)", C);
  NND_REQUIRE(aligned16(x), "flow_head_tail: x must be 16-byte aligned");
  NND_REQUIRE(delta || coords_in, "flow_head_tail: nothing to write (delta and coords_in are both null)");
  NND_REQUIRE(!coords_in || coords_out, "flow_head_tail: coords_in needs coords_out");
  NND_REQUIRE(!flow_out || (coords_in && org), "flow_head_tail: flow_out needs coords_in and org");
  const int strips_per_row = (W + FH_PX - 1) / FH_PX;
  const long long n_strips = static_cast<long long>(N) * H * strips_per_row;
  const long long want = (n_strips + 7) / 8;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define NND_FHT(CPL, F16) \
  flow_head_tail_kernel<CPL, F16><<<grid, 256, 0, st>>>(x, weight, bias, H, W, n_strips, strips_per_row, delta, coords_in, org, \
                                                        coords_out, flow_out)
  if (C == 128) {
    if (x_f16) NND_FHT(4, 1); else NND_FHT(4, 0);
  } else if (C == 256) {
    if (x_f16) NND_FHT(8, 1); else NND_FHT(8, 0);
  } else {
    if (x_f16) NND_FHT(16, 1); else NND_FHT(16, 0);
  }
#undef NND_FHT
  return check_launch("flow_head_tail_kernel");
}

}  // extern "C"
