// CREStereo adaptive group correlation layer (AGCL), sm_100a.
//
//   nnd_agcl_offset   AGCL.corr_att_offset   nndepth/models/cre_stereo/cost_volume.py:81-154
//   nnd_agcl_iter     AGCL.corr_iter + get_correlation   :54-79, :28-52
//   samplers          bilinear_sampler / bilinear_grid_sample   cre_stereo/utils.py:5-20, :34-107
//
// The reference materialises the 9 sampled right maps (N, C/4, 9H, W) per group, repeats the left
// map 9x, multiplies and means; here nothing is materialised -- each output is accumulated in a
// register while the two feature maps are streamed once (the nine taps re-read L1/L2-resident data).
#include "common.cuh"

namespace nnd {

constexpr int AGCL_GROUPS = 4;  // cost_volume.py:69-70, :101-102
constexpr int AGCL_TAPS = 9;    // search_num, cost_volume.py:113

// pixel coordinate -> normalised -> pixel coordinate, in the reference's fp32 operation order:
// utils.py:9-10 `2 * p / (size - 1) - 1`, then utils.py:59-60 `((g + 1) / 2) * (size - 1)`.
__device__ __forceinline__ float pixel_round_trip(float p, float span) {
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, p), span), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), span);
}

// Zero-padded bilinear footprint (utils.py:66-93): four element offsets into one channel plane
// (or -1 when the corner lies outside the image) and the four weights from the UNCLAMPED corners.
struct Footprint {
  int off[4];  // a = (x0,y0), b = (x0,y1), c = (x1,y0), d = (x1,y1)
  float wt[4];
};

__device__ __forceinline__ Footprint make_footprint(float px, float py, int H, int W) {
  const float x = pixel_round_trip(px, static_cast<float>(W - 1));
  const float y = pixel_round_trip(py, static_cast<float>(H - 1));
  const float x0f = floorf(x), y0f = floorf(y);
  const float x1f = __fadd_rn(x0f, 1.0f), y1f = __fadd_rn(y0f, 1.0f);
  Footprint f;
  f.wt[0] = __fmul_rn(__fsub_rn(x1f, x), __fsub_rn(y1f, y));
  f.wt[1] = __fmul_rn(__fsub_rn(x1f, x), __fsub_rn(y, y0f));
  f.wt[2] = __fmul_rn(__fsub_rn(x, x0f), __fsub_rn(y1f, y));
  f.wt[3] = __fmul_rn(__fsub_rn(x, x0f), __fsub_rn(y, y0f));
  // clamp in float first: NaN / huge coordinates become an out-of-image corner (value 0)
  const int x0 = static_cast<int>(fminf(fmaxf(x0f, -2.0f), static_cast<float>(W + 1)));
  const int y0 = static_cast<int>(fminf(fmaxf(y0f, -2.0f), static_cast<float>(H + 1)));
  const int x1 = x0 + 1, y1 = y0 + 1;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
  f.off[0] = (vx0 && vy0) ? y0 * W + x0 : -1;
  f.off[1] = (vx0 && vy1) ? y1 * W + x0 : -1;
  f.off[2] = (vx1 && vy0) ? y0 * W + x1 : -1;
  f.off[3] = (vx1 && vy1) ? y1 * W + x1 : -1;
  return f;
}

__device__ __forceinline__ float corner(const float* __restrict__ plane, int off) {
  return off >= 0 ? __ldg(plane + off) : 0.0f;
}

// Ia*wa + Ib*wb + Ic*wc + Id*wd, left to right, every step rounded (utils.py:107)
__device__ __forceinline__ float blend(const float v[4], const float wt[4]) {
  float r = __fmul_rn(v[0], wt[0]);
  r = __fadd_rn(r, __fmul_rn(v[1], wt[1]));
  r = __fadd_rn(r, __fmul_rn(v[2], wt[2]));
  r = __fadd_rn(r, __fmul_rn(v[3], wt[3]));
  return r;
}

__device__ __forceinline__ void tap_delta(int k, bool small_patch, int& dx, int& dy) {
  if (small_patch) {  // 3x3, dy outer / dx inner (cost_volume.py:121-131 meshgrid 'xy'; :43-44)
    dy = k / 3 - 1;
    dx = k % 3 - 1;
  } else {  // 1x9
    dy = 0;
    dx = k - 4;
  }
}

// ------------------------------------------------------------------------------------------------
// Offset mode.  Block = 32 consecutive pixels x 9 taps (one warp per tap).  Each thread owns one
// (pixel, tap): its deformable sample position is fixed for all channels, so footprint and weights
// are computed once and the channel loop is 1 coalesced left load + 4 near-coalesced right gathers
// + blend + FMA, four channels in flight.  Groups are consecutive channel ranges, so the accumulator
// is flushed every C/4 channels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * AGCL_TAPS)
agcl_offset_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ flow,
                   const float* __restrict__ extra, int C, int H, int W, long long n_pix, int small_patch,
                   float* __restrict__ out) {
  const int k = threadIdx.y;
  const long long pix = static_cast<long long>(blockIdx.x) * 32 + threadIdx.x;
  if (pix >= n_pix) return;
  const long long hw = static_cast<long long>(H) * W;
  const long long n = pix / hw;
  const int p = static_cast<int>(pix - n * hw);
  const int y = p / W, x = p - y * W;

  int dx, dy;
  tap_delta(k, small_patch != 0, dx, dy);
  const float* fl = flow + n * 2 * hw + p;
  const float* ex = extra + (n * 2 * AGCL_TAPS + 2 * k) * hw + p;
  // (grid + flow) + (d_k + extra_k), in that association order (cost_volume.py:133-137)
  const float px = __fadd_rn(__fadd_rn(static_cast<float>(x), __ldg(fl)),
                             __fadd_rn(static_cast<float>(dx), __ldg(ex)));
  const float py = __fadd_rn(__fadd_rn(static_cast<float>(y), __ldg(fl + hw)),
                             __fadd_rn(static_cast<float>(dy), __ldg(ex + hw)));
  const Footprint f = make_footprint(px, py, H, W);

  const int cg = C / AGCL_GROUPS;
  const float inv_cnt_div = static_cast<float>(cg);
  const float* lp = L + n * C * hw + p;
  const float* rp = R + n * C * hw;
  float* op = out + (n * AGCL_GROUPS * AGCL_TAPS + k) * hw + p;
  for (int g = 0; g < AGCL_GROUPS; ++g) {
    float acc = 0.f;
    int c = 0;
    for (; c + 4 <= cg; c += 4) {
      float l[4], v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long ch = static_cast<long long>(g) * cg + c + u;
        l[u] = __ldg(lp + ch * hw);
        const float* plane = rp + ch * hw;
#pragma unroll
        for (int q = 0; q < 4; ++q) v[u][q] = corner(plane, f.off[q]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fmaf(l[u], blend(v[u], f.wt), acc);
    }
    for (; c < cg; ++c) {
      const long long ch = static_cast<long long>(g) * cg + c;
      float v[4];
      const float* plane = rp + ch * hw;
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = corner(plane, f.off[q]);
      acc = fmaf(__ldg(lp + ch * hw), blend(v, f.wt), acc);
    }
    op[static_cast<long long>(g) * AGCL_TAPS * hw] = __fdiv_rn(acc, inv_cnt_div);  // torch.mean over C/4
  }
}

// ------------------------------------------------------------------------------------------------
// Iter mode.  out[n, g*9+k, p] = mean_c L[c,p] * Rw[c, clamp(p + d_k)] with Rw = R warped by the
// flow (zero-padded bilinear at grid + flow) and the clamp = replicate padding of Rw.
// A 256-thread block owns a halo'd tile of the warped map: thread t warps position t of the tile
// (footprint computed once), eight channels at a time go through shared memory, and the interior
// threads accumulate their nine taps from the tile.  The warped map never reaches global memory.
//   1x9 window: tile 64 x 4 positions, interior 56 x 4 outputs (halo 4 columns each side)
//   3x3 window: tile 32 x 8 positions, interior 30 x 6 outputs (halo 1 each side)
// ------------------------------------------------------------------------------------------------
constexpr int IT_CH = 8;

template <bool SMALL>
__global__ void __launch_bounds__(256)
agcl_iter_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ flow, int C,
                 int H, int W, float* __restrict__ out) {
  constexpr int TW = SMALL ? 32 : 64, TH = SMALL ? 8 : 4;
  constexpr int PX = SMALL ? 1 : 4, PY = SMALL ? 1 : 0;
  constexpr int OW = TW - 2 * PX, OH = TH - 2 * PY;
  __shared__ float tile[IT_CH][TH][TW];

  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int x_org = blockIdx.x * OW - PX, y_org = blockIdx.y * OH - PY;  // image position of tile (0,0)
  const int x = x_org + tx, y = y_org + ty;
  const long long n = blockIdx.z;
  const long long hw = static_cast<long long>(H) * W;
  const bool in_image = x >= 0 && x < W && y >= 0 && y < H;
  const bool interior = in_image && tx >= PX && tx < TW - PX && ty >= PY && ty < TH - PY;
  const int p = in_image ? y * W + x : 0;

  Footprint f;
#pragma unroll
  for (int q = 0; q < 4; ++q) { f.off[q] = -1; f.wt[q] = 0.f; }
  if (in_image) {
    const float* fl = flow + n * 2 * hw + p;
    f = make_footprint(__fadd_rn(static_cast<float>(x), __ldg(fl)), __fadd_rn(static_cast<float>(y), __ldg(fl + hw)),
                       H, W);
  }
  // tile-relative word offsets of my nine taps, replicate-clamped into the image
  int tap[AGCL_TAPS];
#pragma unroll
  for (int k = 0; k < AGCL_TAPS; ++k) {
    int dx, dy;
    tap_delta(k, SMALL, dx, dy);
    const int xx = min(max(x + dx, 0), W - 1) - x_org;
    const int yy = min(max(y + dy, 0), H - 1) - y_org;
    tap[k] = interior ? yy * TW + xx : 0;
  }

  const int cg = C / AGCL_GROUPS;
  const float cnt = static_cast<float>(cg);
  const float* lp = L + n * C * hw + p;
  const float* rp = R + n * C * hw;
  float* op = out + n * AGCL_GROUPS * AGCL_TAPS * hw + p;
  float* my_slot = &tile[0][ty][tx];

  for (int g = 0; g < AGCL_GROUPS; ++g) {
    float acc[AGCL_TAPS];
#pragma unroll
    for (int k = 0; k < AGCL_TAPS; ++k) acc[k] = 0.f;
    for (int c0 = 0; c0 < cg; c0 += IT_CH) {
      const int cn = min(IT_CH, cg - c0);
      float warped[IT_CH], l[IT_CH];
#pragma unroll
      for (int u = 0; u < IT_CH; ++u) {
        warped[u] = 0.f;
        l[u] = 0.f;
        if (u < cn) {
          const long long ch = static_cast<long long>(g) * cg + c0 + u;
          if (in_image) {
            const float* plane = rp + ch * hw;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = corner(plane, f.off[q]);
            warped[u] = blend(v, f.wt);
          }
          if (interior) l[u] = __ldg(lp + ch * hw);
        }
      }
      __syncthreads();  // previous chunk's readers are done
#pragma unroll
      for (int u = 0; u < IT_CH; ++u) my_slot[u * TH * TW] = warped[u];
      __syncthreads();
      if (interior) {
#pragma unroll
        for (int u = 0; u < IT_CH; ++u) {
          const float* t = &tile[u][0][0];
#pragma unroll
          for (int k = 0; k < AGCL_TAPS; ++k) acc[k] = fmaf(l[u], t[tap[k]], acc[k]);
        }
      }
    }
    if (interior) {
#pragma unroll
      for (int k = 0; k < AGCL_TAPS; ++k)
        op[(static_cast<long long>(g) * AGCL_TAPS + k) * hw] = __fdiv_rn(acc[k], cnt);
    }
  }
}

static nnd_status check_agcl(const float* f1, const float* f2, const float* flow, const float* out, int N, int C,
                             int H, int W, const char* who) {
  NND_REQUIRE(f1 && f2 && flow && out, "%s: null pointer argument", who);
  NND_REQUIRE(N > 0 && C > 0, "%s: N and C must be positive", who);
  // the samplers divide by (W - 1) and (H - 1) (cre_stereo/utils.py:9-10)
  NND_REQUIRE(H >= 2 && W >= 2, "%s: H and W must be >= 2 (got %d x %d)", who, H, W);
  NND_REQUIRE(C % AGCL_GROUPS == 0, "%s: C = %d is not divisible by the 4 channel groups", who, C);
  NND_REQUIRE(static_cast<long long>(H) * W < (1LL << 30), "%s: feature map too large", who);
  return NND_OK;
}

}  // namespace nnd

extern "C" {

nnd_status nnd_agcl_offset(const float* fmap1, const float* fmap2, const float* flow, const float* extra_offset,
                           int N, int C, int H, int W, int small_patch, float* out, nnd_stream_t stream) {
  using namespace nnd;
  nnd_status st = check_agcl(fmap1, fmap2, flow, out, N, C, H, W, "agcl_offset");
  if (st != NND_OK) return st;
  NND_REQUIRE(extra_offset, "agcl_offset: extra_offset is null");
  const long long n_pix = static_cast<long long>(N) * H * W;
  const long long blocks = (n_pix + 31) / 32;
  NND_REQUIRE(blocks <= 0x7fffffffLL, "agcl_offset: too many pixels");
  dim3 block(32, AGCL_TAPS);
  agcl_offset_kernel<<<static_cast<unsigned>(blocks), block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      fmap1, fmap2, flow, extra_offset, C, H, W, n_pix, small_patch ? 1 : 0, out);
  return check_launch("agcl_offset_kernel");
}

nnd_status nnd_agcl_iter(const float* fmap1, const float* fmap2, const float* flow, int N, int C, int H, int W,
                         int small_patch, float* out, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  nnd_status st = check_agcl(fmap1, fmap2, flow, out, N, C, H, W, "agcl_iter");
  if (st != NND_OK) return st;
  NND_REQUIRE(N <= 65535, "agcl_iter: batch %d exceeds grid limit", N);
  if (small_patch) {
    dim3 grid((W + 29) / 30, (H + 5) / 6, N);
    NND_REQUIRE(grid.y <= 65535, "agcl_iter: feature map too tall");
    agcl_iter_kernel<true><<<grid, 256, 0, stream>>>(fmap1, fmap2, flow, C, H, W, out);
  } else {
    dim3 grid((W + 55) / 56, (H + 3) / 4, N);
    NND_REQUIRE(grid.y <= 65535, "agcl_iter: feature map too tall");
    agcl_iter_kernel<false><<<grid, 256, 0, stream>>>(fmap1, fmap2, flow, C, H, W, out);
  }
  return check_launch("agcl_iter_kernel");
}

}  // extern "C"
