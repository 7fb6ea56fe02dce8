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

// ================================================================================================
// Channels-last fast path (C % 16 == 0).
//
// In NCHW one bilinear corner of one channel is a lone 4-byte gather; with a non-smooth flow a warp
// of 32 pixels touches 32 sectors per corner per channel.  In channels-last (N,H,W,C) a corner is C
// contiguous floats, so every fetched sector is fully used and the loads are 16 bytes per lane.
// The maps are fixed for the life of an AGCL object (one per cascade scale, called 6-12 times), so the
// caller transposes them once (nnd_nchw_to_nhwc) and every call runs on the staged copies.
//
// Work decomposition: one warp per pixel, lanes over channels.  Lane = (group gl = lane / 8, sub-lane
// sl = lane % 8); the 8 sub-lanes of a group stride over that group's C/4 channels in float4 chunks,
// so a warp load is four full 128-byte segments and the group dot product is an 8-lane butterfly.
// A block is 8 warps x 4 pixels = 32 consecutive pixels; the 36 outputs per pixel are parked in
// shared memory and leave as 128-byte coalesced channel-plane stores.
// ================================================================================================
constexpr int CL_WARPS = 8;
constexpr int CL_PIX = 32;           // pixels per block
constexpr int CL_MAX_CHUNKS = 4;     // float4 chunks per lane: C/4 groups of <= 8*4*4 = 128 channels -> C <= 512

// (N,C,H,W) -> (N,H,W,C): 32(c) x 32(hw) tiles through padded shared memory, both sides coalesced.
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, int C, long long hw, float* __restrict__ dst) {
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
    if (c < C && p < hw) dst[(n * hw + p) * C + c] = tile[tx][ty + 8 * i];
  }
}

struct WarpFootprint {  // one tap's bilinear footprint, broadcast to the warp through shared memory
  int off[4];
  float wt[4];
};

__device__ __forceinline__ float4 ld4_or_zero(const float* __restrict__ base, int pix_off, int C, int ch) {
  // pix_off: pixel index inside the image (y*W + x) or -1 for an out-of-image corner
  return pix_off >= 0 ? ldg_f4(base + static_cast<long long>(pix_off) * C + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// wa*a + wb*b + wc*c + wd*d per component (utils.py:107)
__device__ __forceinline__ float4 blend4(const float4 v[4], const float w[4]) {
  float4 r;
  r.x = fmaf(v[3].x, w[3], fmaf(v[2].x, w[2], fmaf(v[1].x, w[1], v[0].x * w[0])));
  r.y = fmaf(v[3].y, w[3], fmaf(v[2].y, w[2], fmaf(v[1].y, w[1], v[0].y * w[0])));
  r.z = fmaf(v[3].z, w[3], fmaf(v[2].z, w[2], fmaf(v[1].z, w[1], v[0].z * w[0])));
  r.w = fmaf(v[3].w, w[3], fmaf(v[2].w, w[2], fmaf(v[1].w, w[1], v[0].w * w[0])));
  return r;
}

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}

__device__ __forceinline__ float group_reduce8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// coalesced write-out of the block's [36][CL_PIX] result tile
__device__ __forceinline__ void store_result_tile(const float (*res)[CL_PIX + 1], long long pix0, long long n_pix,
                                                  long long hw, float* __restrict__ out) {
  const int px = threadIdx.x & 31;
  const long long pix = pix0 + px;
  if (pix >= n_pix) return;
  const long long n = pix / hw, p = pix - n * hw;
  float* op = out + n * AGCL_GROUPS * AGCL_TAPS * hw + p;
  for (int ch = threadIdx.x >> 5; ch < AGCL_GROUPS * AGCL_TAPS; ch += CL_WARPS) op[ch * hw] = res[ch][px];
}

// MODE 0: offset mode (deformable taps, zero-padded bilinear on R)
// MODE 1: iter-mode pass 2 (R = flow-warped map; integer taps, replicate clamp)
template <int MODE>
__global__ void __launch_bounds__(32 * CL_WARPS)
agcl_cl_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ flow,
               const float* __restrict__ extra, int C, int H, int W, long long n_pix, int small_patch,
               float* __restrict__ out) {
  __shared__ WarpFootprint fp[CL_WARPS][AGCL_TAPS];
  __shared__ float res[AGCL_GROUPS * AGCL_TAPS][CL_PIX + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane >> 3, sl = lane & 7;
  const int cg = C / AGCL_GROUPS;
  const int n_chunks = (cg / 4 + 7) / 8;  // float4 chunks per lane
  const long long hw = static_cast<long long>(H) * W;
  const long long pix0 = static_cast<long long>(blockIdx.x) * CL_PIX;
  const float inv_cnt_div = static_cast<float>(cg);

  for (int i = 0; i < CL_PIX / CL_WARPS; ++i) {
    const int slot = warp * (CL_PIX / CL_WARPS) + i;
    const long long pix = pix0 + slot;
    if (pix >= n_pix) break;  // warp-uniform
    const long long n = pix / hw;
    const int p = static_cast<int>(pix - n * hw);
    const int y = p / W, x = p - y * W;

    // taps: lane k < 9 prepares tap k, then everybody reads it back
    if (lane < AGCL_TAPS) {
      int dx, dy;
      tap_delta(lane, small_patch != 0, dx, dy);
      WarpFootprint f;
      if (MODE == 0) {
        const float* fl = flow + n * 2 * hw + p;
        const float* ex = extra + (n * 2 * AGCL_TAPS + 2 * lane) * hw + p;
        // (grid + flow) + (d_k + extra_k), in that association order (cost_volume.py:133-137)
        const float px = __fadd_rn(__fadd_rn(static_cast<float>(x), __ldg(fl)),
                                   __fadd_rn(static_cast<float>(dx), __ldg(ex)));
        const float py = __fadd_rn(__fadd_rn(static_cast<float>(y), __ldg(fl + hw)),
                                   __fadd_rn(static_cast<float>(dy), __ldg(ex + hw)));
        const Footprint ff = make_footprint(px, py, H, W);
        // an out-of-image corner contributes exactly zero (zero padding): give it weight 0 and point it at
        // pixel 0, so the channel loop is predicate-free (float offsets from the image base)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f.off[q] = ff.off[q] >= 0 ? ff.off[q] * C : 0;
          f.wt[q] = ff.off[q] >= 0 ? ff.wt[q] : 0.f;
        }
      } else {
        // replicate padding of the warped map (cost_volume.py:40, utils.py:29-31): clamp the tap into the image
        const int xx = min(max(x + dx, 0), W - 1), yy = min(max(y + dy, 0), H - 1);
        f.off[0] = (yy * W + xx) * C;
        f.off[1] = f.off[2] = f.off[3] = 0;
        f.wt[0] = 1.f;
        f.wt[1] = f.wt[2] = f.wt[3] = 0.f;
      }
      fp[warp][lane] = f;
    }
    __syncwarp();

    const float* lp = L + (n * hw + p) * C + gl * cg + 4 * sl;
    const float* rb = R + n * hw * C + gl * cg + 4 * sl;
    float4 lv[CL_MAX_CHUNKS];
#pragma unroll
    for (int j = 0; j < CL_MAX_CHUNKS; ++j)
      lv[j] = (j < n_chunks && 4 * (sl + 8 * j) < cg) ? ldg_f4(lp + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll(MODE == 1 ? 9 : 3)
    for (int k = 0; k < AGCL_TAPS; ++k) {
      const WarpFootprint f = fp[warp][k];
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < CL_MAX_CHUNKS; ++j) {
        if (j < n_chunks && 4 * (sl + 8 * j) < cg) {
          if (MODE == 0) {
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = ldg_f4(rb + f.off[q] + 32 * j);
            acc = dot4(lv[j], blend4(v, f.wt), acc);
          } else {
            acc = dot4(lv[j], ldg_f4(rb + f.off[0] + 32 * j), acc);
          }
        }
      }
      acc = group_reduce8(acc);
      if (sl == 0) res[gl * AGCL_TAPS + k][slot] = __fdiv_rn(acc, inv_cnt_div);  // torch.mean over C/4
    }
    __syncwarp();
  }
  __syncthreads();
  store_result_tile(res, pix0, n_pix, hw, out);
}

// Four pixels per warp (C = 128 or 256): lane = (pixel = lane / 8, sub-lane sl = lane % 8).  The per-tap
// bookkeeping (footprint fetch, corner addresses, reduction, result store) is the same number of warp
// instructions as in the one-pixel-per-warp kernel above but now serves four pixels, and it was two thirds
// of that kernel's instruction stream (ncu: 1 280 warp instructions per pixel, IPC 2.0, issue-bound -- a
// smooth flow field instead of white noise did not change its time).  The 8 sub-lanes of a pixel read one
// full 128-byte line per chunk; chunk j covers channels 32j .. 32j+31, so it belongs to group j / (CH/4)
// at compile time and each lane keeps four group accumulators.
template <int MODE, int CH>   // CH = C / 32 chunks per lane (4 or 8)
__global__ void __launch_bounds__(32 * CL_WARPS, 2)
agcl_cl4_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ flow,
                const float* __restrict__ extra, int H, int W, long long n_pix, int small_patch, float* __restrict__ out) {
  constexpr int C = 32 * CH;
  constexpr int CPG = CH / AGCL_GROUPS;   // chunks per group
  __shared__ WarpFootprint fp[CL_WARPS][4][AGCL_TAPS];
  __shared__ float res[AGCL_GROUPS * AGCL_TAPS][CL_PIX + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int px = lane >> 3, sl = lane & 7;
  const long long hw = static_cast<long long>(H) * W;
  const long long pix0 = static_cast<long long>(blockIdx.x) * CL_PIX;
  const float inv_cnt_div = static_cast<float>(C / AGCL_GROUPS);

  // footprints of this warp's 4 pixels x 9 taps: 36 (pixel, tap) pairs over 32 lanes, two rounds
  for (int pair = lane; pair < 4 * AGCL_TAPS; pair += 32) {
    const int pp = pair / AGCL_TAPS, k = pair - pp * AGCL_TAPS;
    const long long pix = pix0 + warp * 4 + pp;
    WarpFootprint f;
#pragma unroll
    for (int q = 0; q < 4; ++q) { f.off[q] = 0; f.wt[q] = 0.f; }
    if (pix < n_pix) {
      const long long n = pix / hw;
      const int p = static_cast<int>(pix - n * hw);
      const int y = p / W, x = p - y * W;
      int dx, dy;
      tap_delta(k, small_patch != 0, dx, dy);
      if (MODE == 0) {
        const float* fl = flow + n * 2 * hw + p;
        const float* ex = extra + (n * 2 * AGCL_TAPS + 2 * k) * hw + p;
        // (grid + flow) + (d_k + extra_k), in that association order (cost_volume.py:133-137)
        const float pxf = __fadd_rn(__fadd_rn(static_cast<float>(x), __ldg(fl)), __fadd_rn(static_cast<float>(dx), __ldg(ex)));
        const float pyf = __fadd_rn(__fadd_rn(static_cast<float>(y), __ldg(fl + hw)), __fadd_rn(static_cast<float>(dy), __ldg(ex + hw)));
        const Footprint ff = make_footprint(pxf, pyf, H, W);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f.off[q] = ff.off[q] >= 0 ? ff.off[q] * C : 0;     // out-of-image corner: weight 0, any valid address
          f.wt[q] = ff.off[q] >= 0 ? ff.wt[q] : 0.f;
        }
      } else {
        const int xx = min(max(x + dx, 0), W - 1), yy = min(max(y + dy, 0), H - 1);
        f.off[0] = (yy * W + xx) * C;
        f.wt[0] = 1.f;
      }
    }
    fp[warp][pp][k] = f;
  }
  __syncwarp();

  const int slot = warp * 4 + px;
  const long long pix = pix0 + slot;
  const bool valid = pix < n_pix;
  const long long n = valid ? pix / hw : 0;
  const long long p = valid ? pix - n * hw : 0;
  const float* lp = L + (n * hw + p) * C + 4 * sl;
  const float* rb = R + n * hw * C + 4 * sl;
  float4 lv[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) lv[j] = valid ? ldg_f4(lp + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll 1
  for (int k = 0; k < AGCL_TAPS; ++k) {
    const WarpFootprint f = fp[warp][px][k];
    float acc[AGCL_GROUPS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (MODE == 0) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = ldg_f4(rb + f.off[q] + 32 * j);
        acc[j / CPG] = dot4(lv[j], blend4(v, f.wt), acc[j / CPG]);
      } else {
        acc[j / CPG] = dot4(lv[j], ldg_f4(rb + f.off[0] + 32 * j), acc[j / CPG]);
      }
    }
#pragma unroll
    for (int g = 0; g < AGCL_GROUPS; ++g) {
      const float r = group_reduce8(acc[g]);
      if (sl == 0) res[g * AGCL_TAPS + k][slot] = __fdiv_rn(r, inv_cnt_div);  // torch.mean over C/4
    }
  }
  __syncthreads();
  store_result_tile(res, pix0, n_pix, hw, out);
}

// iter-mode pass 1: Rw[n,p,:] = zero-padded bilinear sample of R at p + flow(p), channels-last in and out
__global__ void __launch_bounds__(32 * CL_WARPS)
agcl_warp_cl_kernel(const float* __restrict__ R, const float* __restrict__ flow, int C, int H, int W, long long n_pix,
                    float* __restrict__ Rw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long hw = static_cast<long long>(H) * W;
  const long long pix = static_cast<long long>(blockIdx.x) * CL_WARPS + warp;
  if (pix >= n_pix) return;
  const long long n = pix / hw;
  const int p = static_cast<int>(pix - n * hw);
  const int y = p / W, x = p - y * W;
  const float* fl = flow + n * 2 * hw + p;
  const Footprint f = make_footprint(__fadd_rn(static_cast<float>(x), __ldg(fl)),
                                     __fadd_rn(static_cast<float>(y), __ldg(fl + hw)), H, W);
  const float* rb = R + n * hw * C;
  float* dst = Rw + pix * C;
  for (int ch = 4 * lane; ch < C; ch += 128) {
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = ld4_or_zero(rb, f.off[q], C, ch);
    // Ia*wa + Ib*wb + Ic*wc + Id*wd, left to right, every step rounded (utils.py:107): the warped map is
    // an intermediate the reference materialises, so it is reproduced bit for bit
    float4 r;
    const float vx[4] = {v[0].x, v[1].x, v[2].x, v[3].x}, vy[4] = {v[0].y, v[1].y, v[2].y, v[3].y};
    const float vz[4] = {v[0].z, v[1].z, v[2].z, v[3].z}, vw[4] = {v[0].w, v[1].w, v[2].w, v[3].w};
    r.x = blend(vx, f.wt); r.y = blend(vy, f.wt); r.z = blend(vz, f.wt); r.w = blend(vw, f.wt);
    *reinterpret_cast<float4*>(dst + ch) = r;
  }
}

// ------------------------------------------------------------------------------------------------
// Iter mode in ONE pass (cost_volume.py:54-79): the flow-warped right map is never written to HBM.  A block owns a
// tile of TH x TW output pixels.  Phase 1 computes the warped map Rw[q] = bilinear(R, q + flow(q)) (zero padding,
// utils.py:34-107, every step rounded like the reference's materialised intermediate) for the tile PLUS the halo its
// nine replicate-clamped taps can reach -- +-4 columns for the 1x9 window, +-1 row / column for 3x3 -- into shared
// memory: each warped pixel is produced once per tile (1.1x / 1.9x the tile's own pixels) instead of being gathered
// nine times from L2.  Phase 2: warp = pixel, lane = C/32 consecutive channels (8 lanes per channel group); the left
// vector comes straight from global memory (1 KB coalesced), each tap is a conflict-free read of the staged tile, an
// 8-lane butterfly finishes the group dot.  Phase 3 writes the [36][TH*TW] result tile as coalesced NCHW row segments.
// Compulsory traffic: each map once (2*C*4 B/px) + flow + output = 2 200 B/px at C = 256.
// ------------------------------------------------------------------------------------------------
// Sum N values over the N consecutive lanes of a group with N - 1 exchanges instead of N * log2(N): at the step with
// distance m every lane passes on the half of its values its partner is responsible for.  Afterwards v[0] of lane l is
// the group's sum of value (l & (N - 1)).
template <int N>
__device__ __forceinline__ float transpose_reduce(float (&v)[N], int lane) {
#pragma unroll
  for (int m = 1, cnt = N / 2; cnt >= 1; m <<= 1, cnt >>= 1) {
    const bool hi = lane & m;
#pragma unroll
    for (int t = 0; t < cnt; ++t) {
      const float send = hi ? v[2 * t] : v[2 * t + 1], keep = hi ? v[2 * t + 1] : v[2 * t];
      v[t] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  return v[0];
}

template <bool SMALL>
struct IterTile {
  static constexpr int TH = SMALL ? 3 : 1;
#ifndef NND_AGCL_TW
#define NND_AGCL_TW 40        // 1x9 tiles: 4 per 160-pixel row, 1 440 CTAs = 4.9 waves of 296 (80: 2.4 waves, 62.5 vs 58.4 us)
#endif
#ifndef NND_AGCL_THREADS
#define NND_AGCL_THREADS 512
#endif
#ifndef NND_AGCL_MINB
#define NND_AGCL_MINB 2
#endif

#ifndef NND_AGCL_TW3
#define NND_AGCL_TW3 14       // 3x3 tiles: 3 x 14 (staged 5 x 16); 3 x 16 measured 74.8 us, 3 x 12 69.6, 3 x 14 65.5, 3 x 18 79.9
#endif
  static constexpr int TW = SMALL ? NND_AGCL_TW3 : NND_AGCL_TW;
  static constexpr int HX = SMALL ? 1 : 4;
  static constexpr int HY = SMALL ? 1 : 0;
  static constexpr int SW = TW + 2 * HX;
  static constexpr int SH = TH + 2 * HY;
  static constexpr int THREADS = NND_AGCL_THREADS;
};

template <bool SMALL, int V>   // V float4 per lane: C = 128 * V
__global__ void __launch_bounds__(IterTile<SMALL>::THREADS, NND_AGCL_MINB)
agcl_iter_fused_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ flow, int H, int W,
                       float* __restrict__ out) {
  using T = IterTile<SMALL>;
  constexpr int C = 128 * V;
  constexpr int NW = T::THREADS / 32;
  constexpr int NS = T::SH * T::SW;          // staged (warped) pixels
  constexpr int NP = T::TH * T::TW;          // output pixels
  extern __shared__ __align__(16) float ism[];
  float* rw = ism;                                                        // [NS][C]
  float* res = rw + NS * C;                                               // [36][NP + 1]
  WarpFootprint* fp = reinterpret_cast<WarpFootprint*>(res + AGCL_GROUPS * AGCL_TAPS * (NP + 1));   // [NS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * T::TW, y0 = blockIdx.y * T::TH;
  const long long hw = static_cast<long long>(H) * W;
  const float* fl = flow + static_cast<long long>(n) * 2 * hw;
  // phase 0: footprints of the staged pixels that lie inside the image.  An out-of-image corner contributes exactly
  // zero (zero padding): it gets weight 0 and points at pixel 0, so the gather below is predicate-free.
  for (int s = tid; s < NS; s += T::THREADS) {
    const int sy = s / T::SW, sx = s - sy * T::SW;
    const int qx = x0 - T::HX + sx, qy = y0 - T::HY + sy;
    WarpFootprint f;
#pragma unroll
    for (int q = 0; q < 4; ++q) { f.off[q] = -1; f.wt[q] = 0.f; }
    if (qx >= 0 && qx < W && qy >= 0 && qy < H) {
      const int p = qy * W + qx;
      const Footprint ff = make_footprint(__fadd_rn(static_cast<float>(qx), __ldg(fl + p)),
                                          __fadd_rn(static_cast<float>(qy), __ldg(fl + hw + p)), H, W);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        f.off[q] = ff.off[q] >= 0 ? ff.off[q] * C : 0;       // float offset inside the image (H*W*C < 2^31)
        f.wt[q] = ff.off[q] >= 0 ? ff.wt[q] : 0.f;
      }
    }
    fp[s] = f;
  }
  __syncthreads();

  // the left vector of this warp's first output pixel is put in flight before the barrier, every later one while its
  // predecessor is being correlated: their DRAM latency never sits between two pixels
  auto load_left = [&](int i, float4 (&dst)[V]) {
    const int ty = i / T::TW, tx = i - ty * T::TW;
    const int x = x0 + tx, y = y0 + ty;
    const bool ok = i < NP && x < W && y < H;
    const float* lp = L + (static_cast<long long>(n) * hw + (ok ? y * W + x : 0)) * C + 4 * lane;
#pragma unroll
    for (int j = 0; j < V; ++j) dst[j] = ok ? ldg_f4(lp + 128 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 lv[V], lnext[V];
  // phase 1: warp per staged pixel: 4 V independent 16-byte gathers per lane, blended in the reference's order
  // lane -> float4 number lane + 32 j of a pixel's channel vector: every warp instruction moves 512 contiguous bytes
  const float* rb = R + static_cast<long long>(n) * hw * C + 4 * lane;
  for (int s = warp; s < NS; s += NW) {
    const WarpFootprint f = fp[s];
    if (f.off[0] < 0) continue;                               // outside the image: never read (taps are clamped into it)
#ifdef NND_AGCL_SKIP_STAGE                                  // timing probe: no gathers of the right map
    continue;
#endif
    float4 v[4][V];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < V; ++j) v[q][j] = ldg_f4(rb + f.off[q] + 128 * j);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      // Ia*wa + Ib*wb + Ic*wc + Id*wd, left to right, every step rounded (utils.py:107)
      const float vx[4] = {v[0][j].x, v[1][j].x, v[2][j].x, v[3][j].x}, vy[4] = {v[0][j].y, v[1][j].y, v[2][j].y, v[3][j].y};
      const float vz[4] = {v[0][j].z, v[1][j].z, v[2][j].z, v[3][j].z}, vw[4] = {v[0][j].w, v[1][j].w, v[2][j].w, v[3][j].w};
      *reinterpret_cast<float4*>(rw + s * C + 4 * lane + 128 * j) =
          make_float4(blend(vx, f.wt), blend(vy, f.wt), blend(vz, f.wt), blend(vw, f.wt));
    }
  }
  load_left(warp, lnext);
  __syncthreads();

  // phase 2: warp per output pixel, taps from the staged tile.  Float4 number lane + 32 j belongs to channel group
  // (lane >> 4) + 2 j at C = 256 (16 lanes per group and j) and to group lane >> 3 at C = 128; every lane accumulates
  // its channels for all nine taps, then the lanes of a group reduce taps 0..7 with the transposing butterfly and tap 8
  // with a plain one.  torch.mean over C/4 = 64 or 32 channels is an exact scaling.
  const float inv_cnt = 1.0f / static_cast<float>(C / AGCL_GROUPS);
  constexpr unsigned FULL = 0xffffffffu;
  for (int i = warp; i < NP; i += NW) {
    const int ty = i / T::TW, tx = i - ty * T::TW;
    const int x = x0 + tx, y = y0 + ty;
#pragma unroll
    for (int j = 0; j < V; ++j) lv[j] = lnext[j];
    load_left(i + NW, lnext);
    if (x >= W || y >= H) continue;                          // warp-uniform
#ifdef NND_AGCL_SKIP_TAPS                                   // timing probe: staging + left loads + output only
    if (lv[0].x == 123456.f) res[i] = lv[V - 1].y;
    continue;
#endif
    float acc[V][AGCL_TAPS];
#pragma unroll
    for (int k = 0; k < AGCL_TAPS; ++k) {
      const int dx = SMALL ? (k % 3 - 1) : (k - 4), dy = SMALL ? (k / 3 - 1) : 0;
      // replicate padding of the warped map (cost_volume.py:40, utils.py:29-31): clamp the tap into the image
      const int qx = min(max(x + dx, 0), W - 1), qy = min(max(y + dy, 0), H - 1);
      const float* rp = rw + ((qy - (y0 - T::HY)) * T::SW + (qx - (x0 - T::HX))) * C + 4 * lane;
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j][k] = dot4(lv[j], *reinterpret_cast<const float4*>(rp + 128 * j), 0.f);
    }
    if (V == 2) {
      float v16[16];
#pragma unroll
      for (int k = 0; k < 8; ++k) { v16[k] = acc[0][k]; v16[8 + k] = acc[V - 1][k]; }
      const float mine = transpose_reduce<16>(v16, lane);                 // value (lane & 15): j = bit 3, tap = low 3 bits
      const int j = (lane >> 3) & 1, k = lane & 7, grp = (lane >> 4) + 2 * j;
      res[(grp * AGCL_TAPS + k) * (NP + 1) + i] = mine * inv_cnt;
      // tap 8: lane bit 0 picks j, then a plain reduction over the other three lane bits
      const bool odd = lane & 1;
      float last = (odd ? acc[V - 1][8] : acc[0][8]) + __shfl_xor_sync(FULL, odd ? acc[0][8] : acc[V - 1][8], 1);
      last += __shfl_xor_sync(FULL, last, 2);
      last += __shfl_xor_sync(FULL, last, 4);
      last += __shfl_xor_sync(FULL, last, 8);
      if ((lane & 14) == 0) res[(((lane >> 4) + 2 * (lane & 1)) * AGCL_TAPS + 8) * (NP + 1) + i] = last * inv_cnt;
    } else {
      float v8[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v8[k] = acc[0][k];
      const float mine = transpose_reduce<8>(v8, lane);                   // tap lane & 7 of group lane >> 3
      const float last = group_reduce8(acc[0][8]);
      res[((lane >> 3) * AGCL_TAPS + (lane & 7)) * (NP + 1) + i] = mine * inv_cnt;
      if ((lane & 7) == 0) res[((lane >> 3) * AGCL_TAPS + 8) * (NP + 1) + i] = last * inv_cnt;
    }
  }
  __syncthreads();

  // phase 3: NCHW output, channel = g*9 + k: TW-wide row segments
  float* ob = out + static_cast<long long>(n) * AGCL_GROUPS * AGCL_TAPS * hw;
  for (int idx = tid; idx < AGCL_GROUPS * AGCL_TAPS * NP; idx += T::THREADS) {
    const int ch = idx / NP, i = idx - ch * NP;
    const int ty = i / T::TW, tx = i - ty * T::TW;
    const int x = x0 + tx, y = y0 + ty;
    if (x < W && y < H) ob[ch * hw + static_cast<long long>(y) * W + x] = res[ch * (NP + 1) + i];
  }
}

template <bool SMALL, int V>
static nnd_status launch_iter_fused(const float* L, const float* R, const float* flow, int N, int H, int W, float* out,
                                    cudaStream_t stream) {
  using T = IterTile<SMALL>;
  constexpr int C = 128 * V;
  constexpr size_t smem = (static_cast<size_t>(T::SH * T::SW) * C + AGCL_GROUPS * AGCL_TAPS * (T::TH * T::TW + 1)) * sizeof(float) +
                          static_cast<size_t>(T::SH * T::SW) * sizeof(WarpFootprint);
  static_assert(smem <= 227 * 1024, "iter tile does not fit shared memory");
  cudaError_t e = cudaFuncSetAttribute(agcl_iter_fused_kernel<SMALL, V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_fail(e, "agcl_iter_nhwc: shared-memory attribute");
  dim3 grid((W + T::TW - 1) / T::TW, (H + T::TH - 1) / T::TH, N);
  NND_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "agcl_iter_nhwc: grid too large");
  agcl_iter_fused_kernel<SMALL, V><<<grid, T::THREADS, smem, stream>>>(L, R, flow, H, W, out);
  return check_launch("agcl_iter_fused_kernel");
}

// ------------------------------------------------------------------------------------------------
// Backward of AGCL for the reference's trainers (cre_stereo/cre_trainer.py; the forward path above is what they
// differentiate through: cre_stereo/cost_volume.py:54-154, utils.py:34-107).  Channels-last maps, warp per pixel, lane =
// 16-byte channel chunks; correctness first (float4 atomics for the scattered map gradients), not tuned.
//
//   sample_backward<OFFSET>  per (pixel p, tap k): position (p + flow(p)) + (d_k + extra_k(p)); the upstream gradient
//                            of channel c is u_c = dout[g(c)*9+k, p] / (C/4) * L[p, c]:
//                              dL[p, c]          += dout/(C/4) * Rs_c          (Rs = the bilinear sample)
//                              dR[corner_q, c]   += w_q * u_c                  (atomic)
//                              d position        += sum_c u_c * dRs_c/d(x, y)  -> d extra_k and d flow
//   sample_backward<WARP>    the same arithmetic with ONE tap at p + flow(p) and u_c = dRw[p, c]: the backward of the
//                            flow-warp that precedes iter mode's window (cost_volume.py:57-59)
//   window_backward          iter mode's replicate-clamped window: dL[p] += dout/(C/4) * Rw[q_k], dRw[q_k] += dout/(C/4) * L[p]
//
// The weights of out-of-image corners come from the UNCLAMPED corner coordinates and multiply a zero image value
// (utils.py:79-93), so they contribute to d position but not to dR.  floor() carries no gradient, the normalise /
// denormalise round trip (utils.py:9-10, 59-60) has derivative 1.
// ------------------------------------------------------------------------------------------------
struct GradFootprint {
  int off[4];        // pixel offsets y*W + x of a, b, c, d or -1
  float fx, fy;      // x - x0, y - y0
};

__device__ __forceinline__ GradFootprint make_grad_footprint(float px, float py, int H, int W) {
  const float x = pixel_round_trip(px, static_cast<float>(W - 1));
  const float y = pixel_round_trip(py, static_cast<float>(H - 1));
  const float x0f = floorf(x), y0f = floorf(y);
  GradFootprint f;
  f.fx = x - x0f;
  f.fy = y - y0f;
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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__device__ __forceinline__ void atomic_add4(float* dst, float4 v) {
  atomicAdd(reinterpret_cast<float4*>(dst), v);   // vector red.global.add (sm_90+)
}

// MODE 0: offset mode (9 taps, upstream = dout * L);  MODE 1: flow-warp backward (1 tap, upstream = dRw)
template <int MODE>
__global__ void __launch_bounds__(256)
agcl_sample_backward_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ flow,
                            const float* __restrict__ extra, const float* __restrict__ up, int C, int H, int W, long long n_pix,
                            int small_patch, float* __restrict__ dL, float* __restrict__ dR, float* __restrict__ dflow,
                            float* __restrict__ dextra) {
  const int lane = threadIdx.x & 31;
  const long long pix = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= n_pix) return;
  const long long hw = static_cast<long long>(H) * W;
  const long long n = pix / hw;
  const int p = static_cast<int>(pix - n * hw);
  const int y = p / W, x = p - y * W;
  const int cg = C / AGCL_GROUPS;
  const float inv_cg = 1.0f / static_cast<float>(cg);
  const float* fl = flow + n * 2 * hw + p;
  const float fxw = __ldg(fl), fyw = __ldg(fl + hw);
  const float* rb = R + n * hw * C;
  float* drb = dR + n * hw * C;
  const float* lp = L ? L + pix * C : nullptr;
  float dfx = 0.f, dfy = 0.f;
  constexpr int NT = MODE == 0 ? AGCL_TAPS : 1;
  for (int k = 0; k < NT; ++k) {
    float px, py;
    if (MODE == 0) {
      int dx, dy;
      tap_delta(k, small_patch != 0, dx, dy);
      const float* ex = extra + (n * 2 * AGCL_TAPS + 2 * k) * hw + p;
      px = __fadd_rn(__fadd_rn(static_cast<float>(x), fxw), __fadd_rn(static_cast<float>(dx), __ldg(ex)));
      py = __fadd_rn(__fadd_rn(static_cast<float>(y), fyw), __fadd_rn(static_cast<float>(dy), __ldg(ex + hw)));
    } else {
      px = __fadd_rn(static_cast<float>(x), fxw);
      py = __fadd_rn(static_cast<float>(y), fyw);
    }
    const GradFootprint f = make_grad_footprint(px, py, H, W);
    const float wa = (1.f - f.fx) * (1.f - f.fy), wb = (1.f - f.fx) * f.fy, wc = f.fx * (1.f - f.fy), wd = f.fx * f.fy;
    float gx = 0.f, gy = 0.f;
    for (int c4 = lane; c4 < C / 4; c4 += 32) {
      const int ch = 4 * c4;
      float4 u;   // upstream gradient of the sampled vector, channels ch .. ch+3
      float go = 0.f;
      float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) {
        go = __ldg(up + (n * AGCL_GROUPS * AGCL_TAPS + (ch / cg) * AGCL_TAPS + k) * hw + p) * inv_cg;
        l = ldg_f4(lp + ch);
        u = make_float4(go * l.x, go * l.y, go * l.z, go * l.w);
      } else {
        u = ldg_f4(up + pix * C + ch);
      }
      float4 r[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) r[q] = f.off[q] >= 0 ? ldg_f4(rb + static_cast<long long>(f.off[q]) * C + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) {
        float4 rs;   // the sample itself: dL = go * Rs
        rs.x = wa * r[0].x + wb * r[1].x + wc * r[2].x + wd * r[3].x;
        rs.y = wa * r[0].y + wb * r[1].y + wc * r[2].y + wd * r[3].y;
        rs.z = wa * r[0].z + wb * r[1].z + wc * r[2].z + wd * r[3].z;
        rs.w = wa * r[0].w + wb * r[1].w + wc * r[2].w + wd * r[3].w;
        float* dl = dL + pix * C + ch;          // a pixel's dL row belongs to this warp: plain read-modify-write
        float4 acc = k == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<float4*>(dl);
        acc.x += go * rs.x; acc.y += go * rs.y; acc.z += go * rs.z; acc.w += go * rs.w;
        *reinterpret_cast<float4*>(dl) = acc;
      }
      const float wq[4] = {wa, wb, wc, wd};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (f.off[q] >= 0)
          atomic_add4(drb + static_cast<long long>(f.off[q]) * C + ch, make_float4(wq[q] * u.x, wq[q] * u.y, wq[q] * u.z, wq[q] * u.w));
      // dRs/dx = (1-fy) (c - a) + fy (d - b);  dRs/dy = (1-fx) (b - a) + fx (d - c)
      const float ddx_x = (1.f - f.fy) * (r[2].x - r[0].x) + f.fy * (r[3].x - r[1].x);
      const float ddx_y = (1.f - f.fy) * (r[2].y - r[0].y) + f.fy * (r[3].y - r[1].y);
      const float ddx_z = (1.f - f.fy) * (r[2].z - r[0].z) + f.fy * (r[3].z - r[1].z);
      const float ddx_w = (1.f - f.fy) * (r[2].w - r[0].w) + f.fy * (r[3].w - r[1].w);
      const float ddy_x = (1.f - f.fx) * (r[1].x - r[0].x) + f.fx * (r[3].x - r[2].x);
      const float ddy_y = (1.f - f.fx) * (r[1].y - r[0].y) + f.fx * (r[3].y - r[2].y);
      const float ddy_z = (1.f - f.fx) * (r[1].z - r[0].z) + f.fx * (r[3].z - r[2].z);
      const float ddy_w = (1.f - f.fx) * (r[1].w - r[0].w) + f.fx * (r[3].w - r[2].w);
      gx += u.x * ddx_x + u.y * ddx_y + u.z * ddx_z + u.w * ddx_w;
      gy += u.x * ddy_x + u.y * ddy_y + u.z * ddy_z + u.w * ddy_w;
    }
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    if (MODE == 0 && lane == 0 && dextra) {
      dextra[(n * 2 * AGCL_TAPS + 2 * k) * hw + p] = gx;
      dextra[(n * 2 * AGCL_TAPS + 2 * k + 1) * hw + p] = gy;
    }
    dfx += gx;
    dfy += gy;
    __syncwarp();   // the dL row is re-read by the next tap
  }
  if (lane == 0 && dflow) {
    dflow[n * 2 * hw + p] = dfx;
    dflow[n * 2 * hw + hw + p] = dfy;
  }
}

__global__ void __launch_bounds__(256)
agcl_window_backward_kernel(const float* __restrict__ L, const float* __restrict__ Rw, const float* __restrict__ dout, int C,
                            int H, int W, long long n_pix, int small_patch, float* __restrict__ dL, float* __restrict__ dRw) {
  const int lane = threadIdx.x & 31;
  const long long pix = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= n_pix) return;
  const long long hw = static_cast<long long>(H) * W;
  const long long n = pix / hw;
  const int p = static_cast<int>(pix - n * hw);
  const int y = p / W, x = p - y * W;
  const int cg = C / AGCL_GROUPS;
  const float inv_cg = 1.0f / static_cast<float>(cg);
  for (int c4 = lane; c4 < C / 4; c4 += 32) {
    const int ch = 4 * c4;
    const float4 l = ldg_f4(L + pix * C + ch);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < AGCL_TAPS; ++k) {
      int dx, dy;
      tap_delta(k, small_patch != 0, dx, dy);
      const int xx = min(max(x + dx, 0), W - 1), yy = min(max(y + dy, 0), H - 1);
      const long long q = (n * hw + static_cast<long long>(yy) * W + xx) * C + ch;
      const float go = __ldg(dout + (n * AGCL_GROUPS * AGCL_TAPS + (ch / cg) * AGCL_TAPS + k) * hw + p) * inv_cg;
      const float4 r = ldg_f4(Rw + q);
      acc.x += go * r.x; acc.y += go * r.y; acc.z += go * r.z; acc.w += go * r.w;
      if (dRw) atomic_add4(dRw + q, make_float4(go * l.x, go * l.y, go * l.z, go * l.w));
    }
    *reinterpret_cast<float4*>(dL + pix * C + ch) = acc;
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

nnd_status nnd_nchw_to_nhwc(const float* src, int N, int C, int H, int W, float* dst, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(src && dst, "nchw_to_nhwc: null pointer");
  NND_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: N, C, H, W must be positive");
  NND_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, "nchw_to_nhwc: N or C exceeds the grid limit");
  const long long hw = static_cast<long long>(H) * W;
  dim3 grid(static_cast<unsigned>((hw + 31) / 32), (C + 31) / 32, N);
  nchw_to_nhwc_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, C, hw, dst);
  return check_launch("nchw_to_nhwc_kernel");
}

static nnd_status check_agcl_cl(int C, int H, int W, const char* who) {
  using namespace nnd;
  NND_REQUIRE(C % 16 == 0, "%s: the channels-last path needs C %% 16 == 0 (got %d); use the NCHW entry point", who, C);
  NND_REQUIRE(C <= 4 * 8 * 4 * CL_MAX_CHUNKS, "%s: C = %d exceeds the channels-last path's limit (%d)", who, C,
              4 * 8 * 4 * CL_MAX_CHUNKS);
  NND_REQUIRE(static_cast<long long>(H) * W * C < (1LL << 31), "%s: one image of the map exceeds 2^31 floats", who);
  return NND_OK;
}

nnd_status nnd_agcl_offset_nhwc(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* flow,
                                const float* extra_offset, int N, int C, int H, int W, int small_patch, float* out,
                                nnd_stream_t stream) {
  using namespace nnd;
  nnd_status st = check_agcl(fmap1_nhwc, fmap2_nhwc, flow, out, N, C, H, W, "agcl_offset_nhwc");
  if (st != NND_OK) return st;
  st = check_agcl_cl(C, H, W, "agcl_offset_nhwc");
  if (st != NND_OK) return st;
  NND_REQUIRE(extra_offset, "agcl_offset_nhwc: extra_offset is null");
  NND_REQUIRE(aligned16(fmap1_nhwc) && aligned16(fmap2_nhwc), "agcl_offset_nhwc: maps must be 16-byte aligned");
  const long long n_pix = static_cast<long long>(N) * H * W;
  const long long blocks = (n_pix + CL_PIX - 1) / CL_PIX;
  NND_REQUIRE(blocks <= 0x7fffffffLL, "agcl_offset_nhwc: too many pixels");
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  if (C == 256) {
    agcl_cl4_kernel<0, 8><<<static_cast<unsigned>(blocks), 32 * CL_WARPS, 0, st_>>>(fmap1_nhwc, fmap2_nhwc, flow, extra_offset, H, W,
                                                                                n_pix, small_patch ? 1 : 0, out);
  } else if (C == 128) {
    agcl_cl4_kernel<0, 4><<<static_cast<unsigned>(blocks), 32 * CL_WARPS, 0, st_>>>(fmap1_nhwc, fmap2_nhwc, flow, extra_offset, H, W,
                                                                                n_pix, small_patch ? 1 : 0, out);
  } else {
    agcl_cl_kernel<0><<<static_cast<unsigned>(blocks), 32 * CL_WARPS, 0, st_>>>(fmap1_nhwc, fmap2_nhwc, flow, extra_offset, C, H, W,
                                                                            n_pix, small_patch ? 1 : 0, out);
  }
  return check_launch("agcl_cl_kernel<offset>");
}

nnd_status nnd_agcl_iter_nhwc(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* flow, int N, int C,
                              int H, int W, int small_patch, float* warped_ws, float* out, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  nnd_status st = check_agcl(fmap1_nhwc, fmap2_nhwc, flow, out, N, C, H, W, "agcl_iter_nhwc");
  if (st != NND_OK) return st;
  st = check_agcl_cl(C, H, W, "agcl_iter_nhwc");
  if (st != NND_OK) return st;
  NND_REQUIRE(warped_ws, "agcl_iter_nhwc: the N*H*W*C workspace for the warped right map is null");
  NND_REQUIRE(aligned16(fmap1_nhwc) && aligned16(fmap2_nhwc) && aligned16(warped_ws),
              "agcl_iter_nhwc: maps and workspace must be 16-byte aligned");
  if (C == 256 || C == 128) {
    // the model's configurations: one fused pass, the warped map lives in shared memory (the workspace stays unused)
    if (small_patch) return C == 256 ? launch_iter_fused<true, 2>(fmap1_nhwc, fmap2_nhwc, flow, N, H, W, out, stream)
                                     : launch_iter_fused<true, 1>(fmap1_nhwc, fmap2_nhwc, flow, N, H, W, out, stream);
    return C == 256 ? launch_iter_fused<false, 2>(fmap1_nhwc, fmap2_nhwc, flow, N, H, W, out, stream)
                    : launch_iter_fused<false, 1>(fmap1_nhwc, fmap2_nhwc, flow, N, H, W, out, stream);
  }
  const long long n_pix = static_cast<long long>(N) * H * W;
  const long long wblocks = (n_pix + CL_WARPS - 1) / CL_WARPS;
  const long long blocks = (n_pix + CL_PIX - 1) / CL_PIX;
  NND_REQUIRE(wblocks <= 0x7fffffffLL, "agcl_iter_nhwc: too many pixels");
  agcl_warp_cl_kernel<<<static_cast<unsigned>(wblocks), 32 * CL_WARPS, 0, stream>>>(fmap2_nhwc, flow, C, H, W, n_pix,
                                                                                    warped_ws);
  st = check_launch("agcl_warp_cl_kernel");
  if (st != NND_OK) return st;
  if (C == 256) {
    agcl_cl4_kernel<1, 8><<<static_cast<unsigned>(blocks), 32 * CL_WARPS, 0, stream>>>(fmap1_nhwc, warped_ws, flow, nullptr, H, W,
                                                                                   n_pix, small_patch ? 1 : 0, out);
  } else if (C == 128) {
    agcl_cl4_kernel<1, 4><<<static_cast<unsigned>(blocks), 32 * CL_WARPS, 0, stream>>>(fmap1_nhwc, warped_ws, flow, nullptr, H, W,
                                                                                   n_pix, small_patch ? 1 : 0, out);
  } else {
    agcl_cl_kernel<1><<<static_cast<unsigned>(blocks), 32 * CL_WARPS, 0, stream>>>(fmap1_nhwc, warped_ws, flow, nullptr, C, H, W,
                                                                               n_pix, small_patch ? 1 : 0, out);
  }
  return check_launch("agcl_cl_kernel<iter>");
}

nnd_status nnd_agcl_warp_nhwc(const float* fmap2_nhwc, const float* flow, int N, int C, int H, int W, float* warped,
                              nnd_stream_t stream) {
  using namespace nnd;
  nnd_status st = check_agcl(fmap2_nhwc, fmap2_nhwc, flow, warped, N, C, H, W, "agcl_warp_nhwc");
  if (st != NND_OK) return st;
  NND_REQUIRE(C % 4 == 0 && aligned16(fmap2_nhwc) && aligned16(warped), "agcl_warp_nhwc: needs C %% 4 == 0 and 16-byte aligned maps");
  const long long n_pix = static_cast<long long>(N) * H * W;
  const long long blocks = (n_pix + CL_WARPS - 1) / CL_WARPS;
  NND_REQUIRE(blocks <= 0x7fffffffLL, "agcl_warp_nhwc: too many pixels");
  agcl_warp_cl_kernel<<<static_cast<unsigned>(blocks), 32 * CL_WARPS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      fmap2_nhwc, flow, C, H, W, n_pix, warped);
  return check_launch("agcl_warp_cl_kernel");
}

nnd_status nnd_agcl_offset_backward_nhwc(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* flow,
                                         const float* extra_offset, const float* grad_out, int N, int C, int H, int W,
                                         int small_patch, float* d_fmap1, float* d_fmap2, float* d_flow, float* d_extra,
                                         nnd_stream_t stream) {
  using namespace nnd;
  nnd_status st = check_agcl(fmap1_nhwc, fmap2_nhwc, flow, d_fmap1, N, C, H, W, "agcl_offset_backward_nhwc");
  if (st != NND_OK) return st;
  NND_REQUIRE(extra_offset && grad_out && d_fmap2, "agcl_offset_backward_nhwc: null pointer argument");
  NND_REQUIRE(C % 16 == 0, "agcl_offset_backward_nhwc: needs C %% 16 == 0 (got %d)", C);
  NND_REQUIRE(aligned16(fmap1_nhwc) && aligned16(fmap2_nhwc) && aligned16(d_fmap1) && aligned16(d_fmap2),
              "agcl_offset_backward_nhwc: maps must be 16-byte aligned");
  const long long n_pix = static_cast<long long>(N) * H * W;
  const long long blocks = (n_pix + 7) / 8;
  NND_REQUIRE(blocks <= 0x7fffffffLL, "agcl_offset_backward_nhwc: too many pixels");
  // d_fmap2 is accumulated with atomics: the caller zero-fills it
  agcl_sample_backward_kernel<0><<<static_cast<unsigned>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      fmap1_nhwc, fmap2_nhwc, flow, extra_offset, grad_out, C, H, W, n_pix, small_patch ? 1 : 0, d_fmap1, d_fmap2, d_flow, d_extra);
  return check_launch("agcl_sample_backward_kernel<offset>");
}

nnd_status nnd_agcl_iter_backward_nhwc(const float* fmap1_nhwc, const float* fmap2_nhwc, const float* flow,
                                       const float* warped, const float* grad_out, int N, int C, int H, int W,
                                       int small_patch, float* d_fmap1, float* d_fmap2, float* d_flow, float* d_warped_ws,
                                       nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  nnd_status st = check_agcl(fmap1_nhwc, fmap2_nhwc, flow, d_fmap1, N, C, H, W, "agcl_iter_backward_nhwc");
  if (st != NND_OK) return st;
  // d_fmap2 == NULL (then d_warped_ws and d_flow are ignored): the reference's own behaviour -- manual_pad detaches the
  // warped right map (cre_stereo/utils.py:29-31), so iter mode trains the LEFT features only
  const bool left_only = d_fmap2 == nullptr;
  NND_REQUIRE(warped && grad_out && (left_only || d_warped_ws), "agcl_iter_backward_nhwc: null pointer argument");
  NND_REQUIRE(C % 16 == 0, "agcl_iter_backward_nhwc: needs C %% 16 == 0 (got %d)", C);
  NND_REQUIRE(aligned16(fmap1_nhwc) && aligned16(fmap2_nhwc) && aligned16(warped) && aligned16(d_fmap1) &&
                  (left_only || (aligned16(d_fmap2) && aligned16(d_warped_ws))),
              "agcl_iter_backward_nhwc: maps must be 16-byte aligned");
  const long long n_pix = static_cast<long long>(N) * H * W;
  const long long blocks = (n_pix + 7) / 8;
  NND_REQUIRE(blocks <= 0x7fffffffLL, "agcl_iter_backward_nhwc: too many pixels");
  // d_fmap2 and d_warped_ws are accumulated with atomics: the caller zero-fills both
  agcl_window_backward_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(fmap1_nhwc, warped, grad_out, C, H, W, n_pix,
                                                                                 small_patch ? 1 : 0, d_fmap1,
                                                                                 left_only ? nullptr : d_warped_ws);
  st = check_launch("agcl_window_backward_kernel");
  if (st != NND_OK || left_only) return st;
  agcl_sample_backward_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      nullptr, fmap2_nhwc, flow, nullptr, d_warped_ws, C, H, W, n_pix, 0, nullptr, d_fmap2, d_flow, nullptr);
  return check_launch("agcl_sample_backward_kernel<warp>");
}

}  // extern "C"
