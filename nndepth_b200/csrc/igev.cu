// IGEV-Stereo: geometry-volume re-layout + pooling, and the soft-argmin regression (sm_100a).
//
//   nnd_geo_transpose_pool   permute/reshape/avg_pool1d chain   igev_stereo/cost_volume.py:44-52
//   nnd_soft_argmin          F.softmax(dim=1) + regress_disparity igev_stereo/model.py:145, :92-95
#include <float.h>

#include "common.cuh"

namespace nnd {

// ------------------------------------------------------------------------------------------------
// geo (B,G,D,H,W1) -> rows [b][g][h][w1] x D, plus the pooled levels, in one pass.
// Per (b,g,h) the source is a D x W1 matrix with row stride H*W1 and the destination is its
// transpose.  A block moves a 128(d) x 32(w1) tile through shared memory: reads are 128-byte rows
// of w1, writes are float4 runs of d, and the pooled levels come out of the same registers
// (store_row_quad), so the 1.5 GB volume is read once and each pyramid level is written once.
// ------------------------------------------------------------------------------------------------
constexpr int GT_D = 128, GT_W = 32;

__global__ void __launch_bounds__(256)
geo_transpose_pool_kernel(const float* __restrict__ geo, int D, int H, int W1, int w_tiles, int d_tiles,
                          int num_levels, Pyramid pyr, int vec_ok) {
  __shared__ float tile[GT_D][GT_W + 1];
  long long bid = blockIdx.x;
  const int wt = static_cast<int>(bid % w_tiles);
  bid /= w_tiles;
  const int dt = static_cast<int>(bid % d_tiles);
  const long long bgh = bid / d_tiles;  // (b*G + g)*H + h
  const long long bg = bgh / H;
  const int h = static_cast<int>(bgh - bg * H);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w_base = wt * GT_W, d_base = dt * GT_D;
  const long long plane = static_cast<long long>(H) * W1;
  const float* src = geo + bg * D * plane + static_cast<long long>(h) * W1;

  const int w = w_base + lane;
#pragma unroll
  for (int i = 0; i < GT_D / 8; ++i) {
    const int dl = warp + 8 * i;
    const int d = d_base + dl;
    tile[dl][lane] = (d < D && w < W1) ? __ldg(src + d * plane + w) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < GT_W / 8; ++i) {
    const int wl = warp + 8 * i;
    const int wr = w_base + wl;
    float4 v;
    v.x = tile[4 * lane + 0][wl];
    v.y = tile[4 * lane + 1][wl];
    v.z = tile[4 * lane + 2][wl];
    v.w = tile[4 * lane + 3][wl];
    const long long row = bgh * W1 + min(wr, W1 - 1);
    store_row_quad(pyr, num_levels, row, d_base + 4 * lane, v, vec_ok != 0, wr < W1);
  }
}

// ------------------------------------------------------------------------------------------------
// Soft-argmin: out[b,0,h,w] = -sum_d d * softmax_d(z[b,d,h,w]) in ONE pass over z (online softmax).
// z is (B,D,H,W): the softmax axis is strided by H*W, consecutive pixels are contiguous.  A block is
// 32 pixels x SA_SLICES disparity slices: each warp streams its slice with 128-byte loads (4 in
// flight per lane), keeping a running (max, sum, weighted sum); the slices are merged through shared
// memory.  197.8 MB in, 1.2 MB out at the IGEV configuration -- pure HBM streaming.
// ------------------------------------------------------------------------------------------------
constexpr int SA_SLICES = 8;
constexpr float LOG2E = 1.4426950408889634f;

struct SoftState {
  float m, s, ws;  // running max, sum of exp(z - m), sum of d * exp(z - m)
};

__device__ __forceinline__ void soft_push4(SoftState& st, const float z[4], int d0, int step, int n_valid) {
  float mx = st.m;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < n_valid) mx = fmaxf(mx, z[i]);
  const float resc = exp2f((st.m - mx) * LOG2E);
  float s = st.s * resc, ws = st.ws * resc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < n_valid) {
      const float e = exp2f((z[i] - mx) * LOG2E);
      s += e;
      ws = fmaf(static_cast<float>(d0 + i * step), e, ws);
    }
  }
  st.m = mx;
  st.s = s;
  st.ws = ws;
}

__global__ void __launch_bounds__(32 * SA_SLICES)
soft_argmin_kernel(const float* __restrict__ z, int D, long long hw, float* __restrict__ out) {
  __shared__ SoftState part[SA_SLICES][32];
  const int lane = threadIdx.x, slice = threadIdx.y;
  const long long p = static_cast<long long>(blockIdx.x) * 32 + lane;
  const long long b = blockIdx.y;
  const bool valid = p < hw;
  const float* src = z + b * D * hw + (valid ? p : 0);

  SoftState st;
  st.m = -FLT_MAX;
  st.s = 0.f;
  st.ws = 0.f;
  // disparities slice, slice + S, slice + 2S, ... ; four loads in flight per lane
  for (int d = slice; d < D; d += 4 * SA_SLICES) {
    float v[4];
    int n_valid = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int di = d + i * SA_SLICES;
      v[i] = 0.f;
      if (di < D) {
        v[i] = __ldcs(src + di * hw);
        n_valid = i + 1;
      }
    }
    soft_push4(st, v, d, SA_SLICES, n_valid);
  }
  part[slice][lane] = st;
  __syncthreads();
  if (slice == 0 && valid) {
    float mx = part[0][lane].m;
#pragma unroll
    for (int i = 1; i < SA_SLICES; ++i) mx = fmaxf(mx, part[i][lane].m);
    float s = 0.f, ws = 0.f;
#pragma unroll
    for (int i = 0; i < SA_SLICES; ++i) {
      const float resc = exp2f((part[i][lane].m - mx) * LOG2E);
      s = fmaf(part[i][lane].s, resc, s);
      ws = fmaf(part[i][lane].ws, resc, ws);
    }
    out[b * hw + p] = -(ws / s);
  }
}

}  // namespace nnd

extern "C" {

nnd_status nnd_geo_transpose_pool(const float* geo, int B, int G, int D, int H, int W1, int num_levels,
                                  float* const* level, const int* pitch, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NND_REQUIRE(geo, "geo_transpose_pool: null volume");
  NND_REQUIRE(B > 0 && G > 0 && D > 0 && H > 0 && W1 > 0, "geo_transpose_pool: B, G, D, H, W1 must be positive");
  Pyramid pyr;
  bool vec_ok;
  nnd_status st = fill_pyramid(pyr, D, num_levels, level, pitch, vec_ok, "geo_transpose_pool");
  if (st != NND_OK) return st;
  const int w_tiles = (W1 + GT_W - 1) / GT_W;
  const int d_tiles = (D + GT_D - 1) / GT_D;
  const long long bgh = static_cast<long long>(B) * G * H;
  const long long blocks = bgh * w_tiles * d_tiles;
  NND_REQUIRE(blocks <= 0x7fffffffLL, "geo_transpose_pool: volume too large for one launch");
  geo_transpose_pool_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(geo, D, H, W1, w_tiles, d_tiles,
                                                                               num_levels, pyr, vec_ok ? 1 : 0);
  st = check_launch("geo_transpose_pool_kernel");
  if (st != NND_OK) return st;
  return pool_tail(pyr, num_levels, bgh * W1, stream);
}

nnd_status nnd_soft_argmin(const float* z, int B, int D, int H, int W, float* out, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(z && out, "soft_argmin: null pointer");
  NND_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, "soft_argmin: B, D, H, W must be positive");
  NND_REQUIRE(B <= 65535, "soft_argmin: batch %d exceeds grid limit", B);
  const long long hw = static_cast<long long>(H) * W;
  dim3 grid(static_cast<unsigned>((hw + 31) / 32), B);
  dim3 block(32, SA_SLICES);
  soft_argmin_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(z, D, hw, out);
  return check_launch("soft_argmin_kernel");
}

}  // extern "C"
