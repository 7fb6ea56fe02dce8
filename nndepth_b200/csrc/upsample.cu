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

__device__ __forceinline__ float ex2_approx(float x) {  // 2**x, MUFU.EX2 (what __expf uses after its multiply)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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
// A pixel's 9*64 logits are contiguous (2304 bytes fp32, 1152 bytes fp16 when the mask head ran as fp16 convolutions).
// lane = (pixel of the warp's 2 or 4, sub-row i, column group): 16 bytes per neighbour k and lane -- four fp32 or
// eight fp16 logits -- so every warp load covers 512 contiguous bytes and a warp stores whole 64 / 128-byte output row
// segments.  The kernel is ISSUE-bound (ncu: 27.6 M warp instructions, issue slots 73 % busy for 138 MB), so the
// arithmetic per logit is kept to add-bias, max, one FMA into the exp2 argument, MUFU.EX2, add, FMA:
//     softmax_k(s * t_k) = exp2(t_k * S - m * S) / sum,   t = logit + bias,  m = max_k t,  S = s * log2(e)  (s > 0)
// and all index arithmetic is 32-bit (the launcher checks the sizes).
template <int MASK_F16>
__global__ void __launch_bounds__(256, 3)
convex_upsample_nhwc8_kernel(const float* __restrict__ flow, const void* __restrict__ mask_, const float* __restrict__ mask_bias,
                             int H, int W, unsigned n_pix, float mask_scale, float* __restrict__ out) {
  constexpr int RATE = 8;
  constexpr int CPL = MASK_F16 ? 8 : 4;            // output columns per lane
  constexpr int LPP = RATE * RATE / CPL;           // lanes per pixel: 8 (fp16) or 16 (fp32)
  constexpr int PPW = 32 / LPP;                    // pixels per warp: 4 or 2
  constexpr int LPR = RATE / CPL;                  // lanes per sub-row: 1 or 2
  constexpr int UPP = 9 * RATE * RATE / (MASK_F16 ? 8 : 4);   // 16-byte units per pixel
  constexpr int UPK = RATE * RATE / (MASK_F16 ? 8 : 4);       // ... per neighbour k
  // The logits travel global -> shared with cp.async, one 16-byte unit per lane and neighbour, two pixel groups in
  // flight per warp: the NEXT group's 4.6 KB are on their way while this one is computed, and they occupy no registers
  // (every lane reads back only what it copied itself, so the stage needs no barrier, only cp.async.wait_group).
  extern __shared__ uint4 up_stage[];              // [warp][2][9][32]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  uint4* mine = up_stage + wib * (2 * 9 * 32) + lane;
  const int sub = lane / LPP, r = lane % LPP, i = r / LPR, jq = r % LPR;
  const unsigned hw = static_cast<unsigned>(H) * static_cast<unsigned>(W);
  const unsigned warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  const float S = mask_scale * 1.4426950408889634f;
  const float4* bp = mask_bias ? reinterpret_cast<const float4*>(mask_bias + i * RATE + CPL * jq) : nullptr;  // L1-resident
  const uint4* units = reinterpret_cast<const uint4*>(mask_) + (MASK_F16 ? i : 2 * i + jq);

  auto issue = [&](unsigned grp, int buf) {
    const unsigned pix = grp * PPW + sub;
    if (grp * PPW < n_pix && pix < n_pix) {
      const uint4* src = units + static_cast<size_t>(pix) * UPP;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(mine + buf * (9 * 32)));
#pragma unroll
      for (int k = 0; k < 9; ++k)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * 512), "l"(src + k * UPK) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue(warp0, 0);
  int buf = 0;
  for (unsigned grp = warp0; grp * PPW < n_pix; grp += n_warps, buf ^= 1) {
    issue(grp + n_warps, buf ^ 1);
    const unsigned pix = grp * PPW + sub;
    const bool live = pix < n_pix;
    const unsigned n = live ? pix / hw : 0, p = live ? pix - n * hw : 0;
    const int h = static_cast<int>(p / static_cast<unsigned>(W)), w = static_cast<int>(p - static_cast<unsigned>(h) * W);
    const float* fl = flow + static_cast<size_t>(n) * hw;
    float nb[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int hh = h + k / 3 - 1, ww = w + k % 3 - 1;
      nb[k] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __fmul_rn(8.0f, __ldg(fl + hh * W + ww)) : 0.f;
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");       // this group's logits have landed
    if (!live) continue;
    const uint4* st = mine + buf * (9 * 32);
    float* op = out + (static_cast<size_t>(n) * RATE * H + static_cast<size_t>(RATE) * h + i) * (static_cast<size_t>(RATE) * W) +
                static_cast<size_t>(RATE) * w + CPL * jq;
#pragma unroll
    for (int g4 = 0; g4 < CPL / 4; ++g4) {
      float4 t[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (MASK_F16) {
          const uint2 hq = *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(st + k * 32) + 8 * g4);
          t[k] = unpack_h4(hq);
        } else {
          const uint4 q = st[k * 32];
          t[k] = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
        }
        if (bp) {
          const float4 bk = __ldg(bp + k * (RATE * RATE / 4) + g4);
          t[k].x += bk.x; t[k].y += bk.y; t[k].z += bk.z; t[k].w += bk.w;
        }
      }
      float res[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = e == 0 ? t[k].x : e == 1 ? t[k].y : e == 2 ? t[k].z : t[k].w;
        float m = v[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) m = fmaxf(m, v[k]);
        const float mS = -m * S;
        float sum = 0.f, a = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float ex = ex2_approx(fmaf(v[k], S, mS));
          sum += ex;
          a = fmaf(ex, nb[k], a);
        }
        res[e] = a / sum;
      }
      *reinterpret_cast<float4*>(op + 4 * g4) = make_float4(res[0], res[1], res[2], res[3]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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
    NND_REQUIRE(n_pix < (1LL << 31) / 64, "convex_upsample: %lld pixels exceed the kernel's 32-bit pixel arithmetic", n_pix);
    NND_REQUIRE(mask_scale > 0.f, "convex_upsample: the channels-last path folds mask_scale into the softmax and needs it positive");
    const int per_block = mask_layout == 2 ? 32 : 16;     // pixels per block of 8 warps
    // one wave of resident blocks (3 per SM by the launch bounds): every warp then loops over ~4 pixel groups and its
    // cp.async pipeline has something to hide
    const long long want = (n_pix + per_block - 1) / per_block, cap = static_cast<long long>(sm_count()) * 3;
    const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
    constexpr int stage_bytes = 8 * 2 * 9 * 32 * 16;      // 8 warps x 2 groups x 9 neighbours x 512 bytes
    if (mask_layout == 2) {
      cudaError_t e = cudaFuncSetAttribute(convex_upsample_nhwc8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage_bytes);
      if (e != cudaSuccess) return cuda_fail(e, "convex_upsample: shared-memory attribute");
      convex_upsample_nhwc8_kernel<1><<<grid, 256, stage_bytes, stream>>>(flow, mask, mask_bias, H, W, static_cast<unsigned>(n_pix), mask_scale, out);
    } else {
      cudaError_t e = cudaFuncSetAttribute(convex_upsample_nhwc8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage_bytes);
      if (e != cudaSuccess) return cuda_fail(e, "convex_upsample: shared-memory attribute");
      convex_upsample_nhwc8_kernel<0><<<grid, 256, stage_bytes, stream>>>(flow, mask, mask_bias, H, W, static_cast<unsigned>(n_pix), mask_scale, out);
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
