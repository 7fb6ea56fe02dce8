// Convex upsampling of the disparity field (sm_100a).
//
//   nnd_convex_upsample   RAFTStereo.convex_upsample   nndepth/models/raft_stereo/model.py:93-105
//                         (same code: cre_stereo/model.py:110-122, igev_stereo/model.py:103-115)
//
// Reference chain per GRU iteration: view -> softmax over the 9 neighbours -> F.unfold(rate * flow, 3x3,
// padding 1) -> multiply -> sum -> permute -> reshape, i.e. five passes over a (N, 9*rate^2, H, W) tensor
// (138 MB at KITTI, batch 8) plus the update block's separate `0.25 *` pass.  Here: one pass.
//   out[n, 0, rate*h + i, rate*w + j] = sum_k softmax_k(s * (mask[n, k*rate^2 + i*rate + j, h, w] + bias[..]))
//                                              * rate * flow[n, 0, h + k/3 - 1, w + k%3 - 1]   (zero padded)
// Thread = (coarse pixel, sub-row i): 9*rate coalesced mask loads (lanes run along w, every load is a full
// 128-byte row segment of one channel plane), a 9-way softmax per output, `rate` consecutive outputs written
// as 16-byte stores.  Pure streaming: 153 MB per launch at the bench shape.
#include <float.h>

#include "common.cuh"

namespace nnd {

template <int RATE>
__global__ void __launch_bounds__(32 * RATE)
convex_upsample_kernel(const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ mask_bias,
                       int H, int W, float mask_scale, float* __restrict__ out) {
  const int lane = threadIdx.x, i = threadIdx.y;
  const long long hw = static_cast<long long>(H) * W;
  const long long p = static_cast<long long>(blockIdx.x) * 32 + lane;
  const long long n = blockIdx.y;
  if (p >= hw) return;
  const int h = static_cast<int>(p / W), w = static_cast<int>(p - static_cast<long long>(h) * W);

  // the 3x3 neighbourhood of rate * flow, zero padded (F.unfold(..., padding=1))
  float nb[9];
  const float* fl = flow + n * hw;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int hh = h + k / 3 - 1, ww = w + k % 3 - 1;
    nb[k] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __fmul_rn(static_cast<float>(RATE), __ldg(fl + static_cast<long long>(hh) * W + ww)) : 0.f;
  }
  const float* mp = mask + (n * 9 * RATE * RATE + i * RATE) * hw + p;
  float res[RATE];
#pragma unroll
  for (int j = 0; j < RATE; ++j) {
    float x[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float bk = mask_bias ? __ldg(mask_bias + k * RATE * RATE + i * RATE + j) : 0.f;
      x[k] = (__ldcs(mp + (static_cast<long long>(k) * RATE * RATE + j) * hw) + bk) * mask_scale;
    }
    float m = x[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) m = fmaxf(m, x[k]);
    float s = 0.f, acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float e = __expf(x[k] - m);
      s += e;
      acc = fmaf(e, nb[k], acc);
    }
    res[j] = acc / s;
  }
  float* op = out + (n * RATE * H + static_cast<long long>(RATE) * h + i) * (static_cast<long long>(RATE) * W) + static_cast<long long>(RATE) * w;
  if (RATE % 4 == 0) {
#pragma unroll
    for (int j = 0; j < RATE; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(res[j], res[j + 1], res[j + 2], res[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < RATE; ++j) op[j] = res[j];
  }
}

// Channels-last mask (N, H, W, 9*RATE*RATE) -- what cuDNN hands back when the hidden state is channels-last.
// Half a warp per coarse pixel: the pixel's 9*64 logits are 2304 contiguous bytes; lane = (pixel of the pair,
// sub-row i, column quad) reads 16 bytes per neighbour k (a warp load covers 2 x 256 contiguous bytes) and
// writes four outputs as one 16-byte store.
// MASK_F16: the logits are IEEE fp16 (the mask head run as fp16 convolutions); a pixel's row is then 1152 bytes.
template <int MASK_F16>
__global__ void __launch_bounds__(256)
convex_upsample_nhwc8_kernel(const float* __restrict__ flow, const void* __restrict__ mask_, const float* __restrict__ mask_bias,
                             int H, int W, long long n_pix, float mask_scale, float* __restrict__ out) {
  constexpr int RATE = 8;
  const int lane = threadIdx.x & 31;
  const int half = lane >> 4, i = (lane >> 1) & 7, jq = lane & 1;
  const long long hw = static_cast<long long>(H) * W;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float4* bp = mask_bias ? reinterpret_cast<const float4*>(mask_bias + i * RATE + 4 * jq) : nullptr;  // L1-resident
  for (long long pair = warp0; 2 * pair < n_pix; pair += n_warps) {
    const long long pix = 2 * pair + half;
    if (pix >= n_pix) continue;
    const long long n = pix / hw;
    const long long p = pix - n * hw;
    const int h = static_cast<int>(p / W), w = static_cast<int>(p - static_cast<long long>(h) * W);
    const float* fl = flow + n * hw;
    const long long quad0 = (pix * (9 * RATE * RATE) + i * RATE + 4 * jq) / 4;   // index in units of 4 logits
    float4 x[9];
    float nb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (MASK_F16) {
        x[k] = unpack_h4(__ldcs(reinterpret_cast<const uint2*>(mask_) + quad0 + k * (RATE * RATE / 4)));
      } else {
        x[k] = __ldcs(reinterpret_cast<const float4*>(mask_) + quad0 + k * (RATE * RATE / 4));
      }
      const int hh = h + k / 3 - 1, ww = w + k % 3 - 1;
      nb[k] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __fmul_rn(8.0f, __ldg(fl + static_cast<long long>(hh) * W + ww)) : 0.f;
    }
    if (bp) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float4 bk = __ldg(bp + k * (RATE * RATE / 4));
        x[k].x += bk.x; x[k].y += bk.y; x[k].z += bk.z; x[k].w += bk.w;
      }
    }
    float res[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = (e == 0 ? x[k].x : e == 1 ? x[k].y : e == 2 ? x[k].z : x[k].w) * mask_scale;
      float m = v[0];
#pragma unroll
      for (int k = 1; k < 9; ++k) m = fmaxf(m, v[k]);
      float s = 0.f, a = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float ex = __expf(v[k] - m);
        s += ex;
        a = fmaf(ex, nb[k], a);
      }
      res[e] = a / s;
    }
    float* op = out + (n * RATE * H + static_cast<long long>(RATE) * h + i) * (static_cast<long long>(RATE) * W) +
                static_cast<long long>(RATE) * w + 4 * jq;
    *reinterpret_cast<float4*>(op) = make_float4(res[0], res[1], res[2], res[3]);
  }
}

}  // namespace nnd

extern "C" {

nnd_status nnd_convex_upsample(const float* flow, const void* mask, const float* mask_bias, int N, int H, int W, int rate,
                               float mask_scale, int mask_layout, float* out, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NND_REQUIRE(flow && mask && out, "convex_upsample: null pointer");
  NND_REQUIRE(N > 0 && H > 0 && W > 0, "convex_upsample: N, H, W must be positive");
  NND_REQUIRE(N <= 65535, "convex_upsample: batch %d exceeds the grid limit", N);
  NND_REQUIRE(rate == 2 || rate == 4 || rate == 8, "convex_upsample: rate %d unsupported (2, 4, 8)", rate);
  NND_REQUIRE(rate % 4 != 0 || aligned16(out), "convex_upsample: output must be 16-byte aligned");
  const long long hw = static_cast<long long>(H) * W;
  NND_REQUIRE(mask_layout >= 0 && mask_layout <= 2,
              "convex_upsample: mask_layout %d is not 0 (fp32 NCHW), 1 (fp32 channels-last) or 2 (fp16 channels-last)", mask_layout);
  if (mask_layout != 0) {
    NND_REQUIRE(rate == 8, "convex_upsample: the channels-last mask path is built for rate 8 (got %d)", rate);
    NND_REQUIRE(aligned16(mask) && aligned16(out) && (!mask_bias || aligned16(mask_bias)),
                "convex_upsample: mask, bias and output must be 16-byte aligned");
    const long long n_pix = hw * N;
    const long long want = (n_pix + 15) / 16, cap = static_cast<long long>(sm_count()) * 8;
    const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
    if (mask_layout == 2) {
      convex_upsample_nhwc8_kernel<1><<<grid, 256, 0, stream>>>(flow, mask, mask_bias, H, W, n_pix, mask_scale, out);
    } else {
      convex_upsample_nhwc8_kernel<0><<<grid, 256, 0, stream>>>(flow, mask, mask_bias, H, W, n_pix, mask_scale, out);
    }
    return check_launch("convex_upsample_nhwc8_kernel");
  }
  dim3 grid(static_cast<unsigned>((hw + 31) / 32), N);
  const float* fmask = reinterpret_cast<const float*>(mask);
  if (rate == 8) {
    convex_upsample_kernel<8><<<grid, dim3(32, 8), 0, stream>>>(flow, fmask, mask_bias, H, W, mask_scale, out);
  } else if (rate == 4) {
    convex_upsample_kernel<4><<<grid, dim3(32, 4), 0, stream>>>(flow, fmask, mask_bias, H, W, mask_scale, out);
  } else {
    convex_upsample_kernel<2><<<grid, dim3(32, 2), 0, stream>>>(flow, fmask, mask_bias, H, W, mask_scale, out);
  }
  return check_launch("convex_upsample_kernel");
}

}  // extern "C"
