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
// 32 lanes x S disparity slices; a lane owns VEC consecutive pixels (one 16-byte load per disparity
// when H*W % 4 == 0), warp `s` streams disparities s, s+S, ... with eight loads in flight per lane and
// keeps a running (max, sum, weighted sum) per pixel; the slices are merged through shared memory.
// 197.8 MB in, 1.2 MB out at the IGEV configuration -- pure HBM streaming.  With enough pixels S = 1:
// every warp is its own block, all of them are resident at once and do equal work, so there is no wave
// quantisation and no merge; small problems slice the disparity axis to fill the machine.
// ------------------------------------------------------------------------------------------------
constexpr int SA_MAX_SLICES = 8;
constexpr float LOG2E = 1.4426950408889634f;

struct SoftState {
  float m, s, ws;  // running max (times log2e), sum of exp(z - max), sum of d * exp(z - max)
};

__device__ __forceinline__ float ex2_fast(float x) {  // 2**x, MUFU.EX2 (2 ulp), flushes denormal results to zero
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Online softmax-expectation over NV more disparities d0, d0+step, ...: one rescale of the running sums
// per call, then exp2(z*log2e - max*log2e) as a single FFMA + MUFU per element.  st.m is kept in the
// log2 domain (max(z) * log2e) so the subtraction folds into the FFMA.
template <int NV>
__device__ __forceinline__ void soft_push(SoftState& st, const float (&z)[NV], int d0, int step, int n_valid) {
  float mx = st.m;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (i < n_valid) mx = fmaxf(mx, z[i] * LOG2E);
  const float resc = ex2_fast(st.m - mx);
  float s = st.s * resc, ws = st.ws * resc;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (i < n_valid) {
      const float e = ex2_fast(fmaf(z[i], LOG2E, -mx));
      s += e;
      ws = fmaf(static_cast<float>(d0 + i * step), e, ws);
    }
  }
  st.m = mx;
  st.s = s;
  st.ws = ws;
}

template <int VEC>
struct SoftVec;
template <>
struct SoftVec<1> {
  using type = float;
  static __device__ __forceinline__ void unpack(float v, float (&o)[1]) { o[0] = v; }
  static __device__ __forceinline__ float pack(const float (&o)[1]) { return o[0]; }
};
template <>
struct SoftVec<4> {
  using type = float4;
  static __device__ __forceinline__ void unpack(float4 v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
  static __device__ __forceinline__ float4 pack(const float (&o)[4]) { return make_float4(o[0], o[1], o[2], o[3]); }
};

// hwv = H*W / VEC pixel groups per image; blockDim = (32, S)
template <int VEC>
__global__ void __launch_bounds__(32 * SA_MAX_SLICES)
soft_argmin_kernel(const float* __restrict__ z, int D, long long hwv, float* __restrict__ out) {
  using V = typename SoftVec<VEC>::type;
  extern __shared__ SoftState part[];  // [S][32][VEC]
  const int lane = threadIdx.x, slice = threadIdx.y, S = blockDim.y;
  const long long p = static_cast<long long>(blockIdx.x) * 32 + lane;
  const long long b = blockIdx.y;
  const bool valid = p < hwv;
  const V* src = reinterpret_cast<const V*>(z) + b * D * hwv + (valid ? p : 0);

  SoftState st[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    st[c].m = -FLT_MAX;
    st[c].s = 0.f;
    st[c].ws = 0.f;
  }
  // disparities slice, slice + S, slice + 2S, ... ; eight loads in flight per lane
  for (int d = slice; d < D; d += 8 * S) {
    float v[8][VEC];
    int n_valid = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) v[i][c] = 0.f;
      if (d + i * S < D) {
        SoftVec<VEC>::unpack(__ldcs(src + (d + i * S) * hwv), v[i]);
        n_valid = i + 1;
      }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      const float zc[8] = {v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]};
      soft_push<8>(st[c], zc, d, S, n_valid);
    }
  }
#pragma unroll
  for (int c = 0; c < VEC; ++c) part[(slice * 32 + lane) * VEC + c] = st[c];
  __syncthreads();
  if (slice == 0 && valid) {
    float res[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      float mx = -FLT_MAX;
      for (int i = 0; i < S; ++i) mx = fmaxf(mx, part[(i * 32 + lane) * VEC + c].m);
      float s = 0.f, ws = 0.f;
      for (int i = 0; i < S; ++i) {
        const SoftState q = part[(i * 32 + lane) * VEC + c];
        const float resc = ex2_fast(q.m - mx);  // m is already in the log2 domain
        s = fmaf(q.s, resc, s);
        ws = fmaf(q.ws, resc, ws);
      }
      res[c] = -(ws / s);
    }
    reinterpret_cast<V*>(out)[b * hwv + p] = SoftVec<VEC>::pack(res);
  }
}

// slices per block: 1 when one warp per 32 pixel groups already fills the machine (>= 12 warps per SM),
// otherwise enough disparity slices to get there (merge cost grows with S, so at most 8).
static int soft_argmin_slices(long long units, int D) {
  const long long want = static_cast<long long>(sm_count()) * 12;
  int S = 1;
  while (S < 8 && units * S < want && 2 * S <= D) S *= 2;
  return S;
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
  const bool vec4 = hw % 4 == 0 && aligned16(z) && aligned16(out);
  const long long hwv = vec4 ? hw / 4 : hw;
  const long long gx = (hwv + 31) / 32;
  NND_REQUIRE(gx <= 0x7fffffffLL, "soft_argmin: H*W too large");
  const int S = soft_argmin_slices(gx * B, D);
  dim3 grid(static_cast<unsigned>(gx), B);
  dim3 block(32, S);
  const size_t smem = static_cast<size_t>(S) * 32 * (vec4 ? 4 : 1) * sizeof(SoftState);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (vec4) {
    soft_argmin_kernel<4><<<grid, block, smem, st>>>(z, D, hwv, out);
  } else {
    soft_argmin_kernel<1><<<grid, block, smem, st>>>(z, D, hwv, out);
  }
  return check_launch("soft_argmin_kernel");
}

}  // extern "C"
