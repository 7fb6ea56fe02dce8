// Fused pyramid lookups (sm_100a): every level, every tap, NCHW output, one launch.
//
//   nnd_corr1d_lookup          CorrBlock1D.__call__            raft_stereo/cost_volume.py:36-53
//   nnd_group_lookup  mode 0   GeometryAwareCostVolume.forward igev_stereo/cost_volume.py:54-79
//                     mode 1   GroupCorrBlock1D.__call__       raft_stereo/cost_volume.py:92-111
//   nnd_corr1d_lookup_indices  the integer half of linear_sampler, raft_stereo/utils.py:16-21
//
// Work decomposition: one warp = 32 consecutive pixels of one image x one pyramid level (x a chunk of
// planes: plane = (source pyramid, group)).  A pixel's window at a level is <= 2r+3 consecutive floats
// of its own volume row, so neighbouring pixels never share data and a thread-per-pixel gather would
// cost one L1 wavefront per lane per tap.  Instead the warp loads the 32 windows cooperatively --
// WINQ lanes per pixel, one 16-byte load each, 32/WINQ rows per instruction -- parks them in a
// padded shared-memory tile and then every lane interpolates its own pixel's taps from shared
// memory.  That is the minimum number of L1 wavefronts (one per touched row segment), the loads of
// a warp are all independent (WINQ in flight per lane), and the stores are 128-byte coalesced
// channel planes.  The kernel is instruction-issue sensitive (a launch moves only ~18 MB at the
// KITTI shape), so the per-tap arithmetic is kept minimal: no 64-bit index math in the loops and an
// exact 3-instruction division (sampler_quotient) instead of the generic IEEE division routine.
//
// Bit-exactness: tap positions follow the reference's fp32 operation order exactly; integer indices
// therefore equal the CPU reference's, and the lerp is evaluated without FMA contraction.
#include <stdlib.h>

#include "common.cuh"
#include "lookup_common.cuh"

namespace nnd {

// WINQ : 16-byte quads per staged window (window = 4*WINQ floats >= 2r+6)
// TAPS : compile-time tap count (2r+1) -> per-tap state lives in registers across planes; 0 = dynamic
template <int WINQ, int TAPS>
__global__ void __launch_bounds__(32 * NND_MAX_LEVELS)
pyramid_lookup_kernel(const LookupArgs a) {
  constexpr int COLS = 4 * WINQ;
  constexpr int STRIDE = COLS + 1;  // odd word stride: lanes reading the same column hit 32 banks
  constexpr int PPL = 32 / WINQ;    // pixels per cooperative load instruction
  constexpr int NT = TAPS > 0 ? TAPS : 1;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ float smem[];

  const int lane = threadIdx.x;
  const int lvl = threadIdx.y;
  float* win = smem + lvl * (STRIDE * 32);

  const int b = blockIdx.z;
  const int rem0 = blockIdx.x * 32;   // first pixel (within the image) of this warp
  const int rem = rem0 + lane;
  const bool valid = rem < a.hw;
  const int r = a.radius;
  const int T = TAPS > 0 ? TAPS : 2 * r + 1;
  const int w = a.src[0].width[lvl];
  const int pitch = a.src[0].pitch[lvl];

  const float c = valid ? __ldg(a.coords + static_cast<long long>(b) * a.hw + rem) : 0.0f;
  // coords / 2**lvl: scaling by a power of two is exact, so the product equals the IEEE quotient
  const float centre = __fmul_rn(c, 1.0f / static_cast<float>(1 << lvl));
  const LevelScale sc = level_scale(w, lvl, centre);

  // window range [lo, hi]: taps are monotone in k, so the extremes come from the first / last tap
  int o0[NT], o1[NT];
  float cf[NT], omc[NT];
  int lo, hi;
  if (TAPS > 0) {
    Tap tp[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) tp[k] = make_tap(k, r, centre, sc);
    lo = tp[0].i0;
    hi = tp[NT - 1].i1;
    const int base = lane * STRIDE - (lo & ~3);
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      o0[k] = base + tp[k].i0;
      o1[k] = base + tp[k].i1;
      cf[k] = tp[k].coef;
      omc[k] = tp[k].one_minus;
    }
  } else {
    lo = make_tap(0, r, centre, sc).i0;
    hi = make_tap(2 * r, r, centre, sc).i1;
  }
  const int s = lo & ~3;  // window start, quad aligned (rows are 16-byte aligned in vec mode)

  // cooperative loader role: this lane fetches quad q of pixel (j*PPL + lane/WINQ), j < WINQ
  const int q = lane % WINQ;
  int goff[WINQ];  // float offset of my quad relative to the warp's first volume row, or -1
  int left[WINQ];  // valid columns from the start of my quad (scalar path)
#pragma unroll
  for (int j = 0; j < WINQ; ++j) {
    const int p = j * PPL + lane / WINQ;
    const int sp = __shfl_sync(FULL, s, p);
    const int hp = __shfl_sync(FULL, hi, p);
    const int cq = sp + 4 * q;
    const bool need = (rem0 + p < a.hw) && cq <= hp && cq < w;
    goff[j] = need ? p * pitch + cq : -1;
    left[j] = w - cq;
  }
  float* const my_quad = win + (lane / WINQ) * STRIDE + 4 * q;  // + j*PPL*STRIDE

  const int GT = a.G * T;
  const int n_planes = a.n_src * a.G;
  const int plane_begin = blockIdx.y * a.planes_per_block;
  const int plane_end = min(plane_begin + a.planes_per_block, n_planes);
  const long long c_total = static_cast<long long>(a.num_levels) * a.n_src * GT;
  float* const out_img = a.out + static_cast<long long>(b) * c_total * a.hw;

  for (int plane = plane_begin; plane < plane_end; ++plane) {
    const int sidx = plane >= a.G ? 1 : 0;
    const int g = plane - sidx * a.G;
    // first volume row of this warp: ((b*G + g) * hw + rem0)
    const float* __restrict__ rows =
        a.src[sidx].ptr[lvl] + ((static_cast<long long>(b) * a.G + g) * a.hw + rem0) * pitch;

    float4 v[WINQ];
#pragma unroll
    for (int j = 0; j < WINQ; ++j) {
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (goff[j] >= 0) {
        const float* p = rows + goff[j];
        if (a.vec) {
          v[j] = ldg_f4(p);
        } else {
          v[j].x = __ldg(p);
          if (left[j] > 1) v[j].y = __ldg(p + 1);
          if (left[j] > 2) v[j].z = __ldg(p + 2);
          if (left[j] > 3) v[j].w = __ldg(p + 3);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < WINQ; ++j) {
      float* dst = my_quad + j * (PPL * STRIDE);
      dst[0] = v[j].x;
      dst[1] = v[j].y;
      dst[2] = v[j].z;
      dst[3] = v[j].w;
    }
    __syncwarp();

    if (valid) {
      if (a.mode == 0) {
        float* op = out_img + (static_cast<long long>(lvl) * a.n_src * GT + sidx * GT + g * T) * a.hw + rem;
        if (TAPS > 0) {
#pragma unroll
          for (int k = 0; k < NT; ++k) {
            // coef * val0 + (1 - coef) * val1, each operation rounded (utils.py:27)
            op[static_cast<long long>(k) * a.hw] = __fadd_rn(__fmul_rn(cf[k], win[o0[k]]), __fmul_rn(omc[k], win[o1[k]]));
          }
        } else {
          const int base = lane * STRIDE - s;
          for (int k = 0; k < T; ++k) {
            const Tap tp = make_tap(k, r, centre, sc);
            const int i0 = min(max(base + tp.i0, lane * STRIDE), lane * STRIDE + COLS - 1);
            const int i1 = min(max(base + tp.i1, lane * STRIDE), lane * STRIDE + COLS - 1);
            op[static_cast<long long>(k) * a.hw] = __fadd_rn(__fmul_rn(tp.coef, win[i0]), __fmul_rn(tp.one_minus, win[i1]));
          }
        }
      } else {
        // GroupCorrBlock1D: the sampled block, in memory order [b][g][h][w][k], is *viewed* as
        // (B, H, W, G*T) (raft_stereo/cost_volume.py:108) and then permuted to NCHW.
        float* op = out_img + static_cast<long long>(lvl) * GT * a.hw;
        const int base = lane * STRIDE - s;
        for (int k = 0; k < T; ++k) {
          const Tap tp = make_tap(k, r, centre, sc);
          const int i0 = min(max(base + tp.i0, lane * STRIDE), lane * STRIDE + COLS - 1);
          const int i1 = min(max(base + tp.i1, lane * STRIDE), lane * STRIDE + COLS - 1);
          const float res = __fadd_rn(__fmul_rn(tp.coef, win[i0]), __fmul_rn(tp.one_minus, win[i1]));
          const long long f = (static_cast<long long>(g) * a.hw + rem) * T + k;
          op[(f % GT) * a.hw + f / GT] = res;
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// RAFT-Stereo fast path (one source, one plane): warp = 32 consecutive pixels x one level, like the
// generic kernel, but lean in registers.  A launch at the KITTI shape moves only ~18 MB, so its
// duration is a latency chain (coords -> window bounds -> window fetch -> interpolate -> store), and the
// generic kernel's 74 registers (per-tap state of 9 taps kept across the plane loop) leave room for only
// half of the 7 488 warps at once: the second wave pays the whole chain again.  Here only the window
// bounds are computed before the fetch and the taps are re-derived afterwards (11 instead of 9 tap
// evaluations), which fits 40 registers -> the whole grid is resident in one wave.
// ------------------------------------------------------------------------------------------------
template <int TAPS>
__global__ void __launch_bounds__(32 * NND_MAX_LEVELS, 6)
corr1d_lookup_lean_kernel(const __grid_constant__ LookupArgs a, unsigned magic_shl2) {
  constexpr int WINQ = 4, COLS = 4 * WINQ, STRIDE = COLS + 1;
  constexpr int R = (TAPS - 1) / 2;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ float smem[];
  const int lane = threadIdx.x, lvl = threadIdx.y;
  float* win = smem + lvl * (32 * STRIDE);

  const int b = blockIdx.y;
  const int rem0 = blockIdx.x * 32;
  const int rem = rem0 + lane;
  const bool valid = rem < a.hw;
  const int w = a.src[0].width[lvl];
  const int pitch = a.src[0].pitch[lvl];
  const float c = valid ? __ldg(a.coords + static_cast<long long>(b) * a.hw + rem) : 0.0f;
  const float centre = __fmul_rn(c, 1.0f / static_cast<float>(1 << lvl));
  const LevelScale sc = level_scale(w, lvl, centre);
  // conservative window: a 4-aligned start never above the first tap's index and at most one below it, an end never
  // below the last tap's (positions differ from centre + dx by a few ulps only; the launcher bounds the width), so the
  // 16 floats fetched hold every tap and no exact tap has to be computed before the fetch
  const int s = __float2int_rd(fminf(fmaxf(__fadd_rn(centre, -4.015625f), 0.f), sc.span)) & ~3;
  const int hi = __float2int_ru(fminf(fmaxf(__fadd_rn(centre, 4.015625f), 0.f), sc.span));

  const int q = lane & 3;
  const float* __restrict__ rows = a.src[0].ptr[lvl] + (static_cast<long long>(b) * a.hw + rem0) * pitch;
  float4 v[WINQ];
#pragma unroll
  for (int j = 0; j < WINQ; ++j) {
    const int p = j * 8 + (lane >> 2);
    const int sp = __shfl_sync(FULL, s, p);
    const int hp = __shfl_sync(FULL, hi, p);
    const int cq = sp + 4 * q;
    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((rem0 + p < a.hw) && cq <= hp && cq < w) {
      const float* src = rows + static_cast<long long>(p) * pitch + cq;
      if (a.vec) {
        v[j] = ldg_f4(src);
      } else {
        const int left = w - cq;
        v[j].x = __ldg(src);
        if (left > 1) v[j].y = __ldg(src + 1);
        if (left > 2) v[j].z = __ldg(src + 2);
        if (left > 3) v[j].w = __ldg(src + 3);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < WINQ; ++j) {
    float* dst = win + (j * 8 + (lane >> 2)) * STRIDE + 4 * q;
    dst[0] = v[j].x;
    dst[1] = v[j].y;
    dst[2] = v[j].z;
    dst[3] = v[j].w;
  }
  __syncwarp();
  if (!valid) return;
  // taps: the reference's operation order (utils.py:16-27) with ONE clamp per pixel (outside [-6, span + 6] every tap
  // saturates to t = 0 or t = span either way; NaN -> -6 -> t = 0) and floor through FADD.RZ with 2^23, whose bits also
  // address the window (magic_shl2 = 0x4b000000 << 2 comes as an argument so that it stays folded into the base)
  const float cc = fminf(fmaxf(centre, -6.0f), __fadd_rn(sc.span, 6.0f));
  const uint32_t mine = static_cast<uint32_t>(__cvta_generic_to_shared(win + lane * STRIDE)) - 4u * static_cast<uint32_t>(s) - magic_shl2;
  float* op = a.out + (static_cast<long long>(b) * a.num_levels + lvl) * TAPS * a.hw + rem;
#pragma unroll
  for (int k = 0; k < TAPS; ++k) {
    const float x = __fadd_rn(static_cast<float>(k - R), cc);
    const float qn = __fmul_rn(x, sc.inv_span);
    const float rr = __fmaf_rn(-qn, sc.span, x);
    const float tt = __fmul_rn(__saturatef(__fmaf_rn(rr, sc.inv_span, qn)), sc.span);
    const float u = __fadd_rz(tt, 8388608.0f);
    const float f0 = __fadd_rn(u, -8388608.0f);
    const uint32_t addr = mine + (__float_as_uint(u) << 2);
    float v0, v1;
    asm volatile("ld.shared.f32 %0, [%2];\n\tld.shared.f32 %1, [%2+4];" : "=f"(v0), "=f"(v1) : "r"(addr) : "memory");
    const bool whole = (tt == f0);
    const float coef = whole ? 0.0f : __fsub_rn(__fadd_rn(f0, 1.0f), tt);   // coef = idx1 - t (utils.py:26)
    const float one_minus = __fsub_rn(1.0f, coef);
    v1 = whole ? v0 : v1;
    // coef * val0 + (1 - coef) * val1, each operation rounded (utils.py:27)
    *op = __fadd_rn(__fmul_rn(coef, v0), __fmul_rn(one_minus, v1));
    op += a.hw;
  }
}

// ------------------------------------------------------------------------------------------------
// Lookup fused with the motion encoder's first layer: out = relu(W . lookup(coords) + b), W = convc1's 1x1
// weights (blocks/update_block.py:51,58: `cor = F.relu(self.convc1(corr))`).  The (B, 36, H, W) lookup
// tensor never reaches HBM: a block computes the 36 taps of 32 pixels into shared memory (one warp per
// level, same arithmetic as the stand-alone kernel), then its four warps turn them into Cout channels in
// fp32 FFMA -- thread = one pixel x Cout/4 output channels, weights read as 16-byte shared-memory
// broadcasts -- and store 128-byte coalesced channel-plane segments.  At the KITTI shape one launch is
// 552 MFMA and a 61 MB write: ~15 us of FFMA, hidden behind nothing else the GPU has to do here.
// Requires num_levels * TAPS % 4 == 0 (36 for RAFT-Stereo) and num_levels == blockDim.y.
// ------------------------------------------------------------------------------------------------
template <int TAPS>
__global__ void __launch_bounds__(128)
corr1d_lookup_conv1x1_kernel(const __grid_constant__ LookupArgs a, const float* __restrict__ weight,
                             const float* __restrict__ bias, int c_out, int relu, long long n_groups) {
  constexpr int WINQ = 4, COLS = 4 * WINQ, STRIDE = COLS + 1;
  constexpr int R = (TAPS - 1) / 2;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float smem[];
  const int L = blockDim.y;
  const int K = L * TAPS;                       // lookup channels (36)
  const int c_pad = (c_out + 3) & ~3;
  float* wsm = smem;                            // [K][c_pad]   k-major: four output channels of one k are 16 bytes
  float* bsm = wsm + K * c_pad;                 // [c_pad]
  float* corr = bsm + c_pad;                    // [K][32]
  float* wins = corr + K * 32;                  // [L][32][STRIDE]
  const int lane = threadIdx.x, lvl = threadIdx.y;
  const int tid = lvl * 32 + lane, nthreads = 32 * L;
  if (c_pad == c_out && (reinterpret_cast<uintptr_t>(weight) & 15u) == 0) {
    // 16-byte copies, eight in flight per thread (a load->store loop would serialise one L2 latency per element)
    const int n4 = K * c_out / 4;
    for (int base = tid; base < n4; base += 8 * nthreads) {
      float4 t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (base + j * nthreads < n4) t[j] = __ldg(reinterpret_cast<const float4*>(weight) + base + j * nthreads);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (base + j * nthreads < n4) reinterpret_cast<float4*>(wsm)[base + j * nthreads] = t[j];
    }
  } else {
    for (int i = tid; i < K * c_pad; i += nthreads) {
      const int k = i / c_pad, co = i - k * c_pad;
      wsm[i] = co < c_out ? __ldg(weight + static_cast<long long>(k) * c_out + co) : 0.f;
    }
  }
  for (int i = tid; i < c_pad; i += nthreads) bsm[i] = (bias && i < c_out) ? __ldg(bias + i) : 0.f;

  float* win = wins + lvl * (32 * STRIDE);
  // persistent blocks: the weights are staged once, pixel groups are dealt round-robin
  const int groups_per_image = (a.hw + 31) / 32;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
  const int b = static_cast<int>(grp / groups_per_image);
  const int rem0 = static_cast<int>(grp - static_cast<long long>(b) * groups_per_image) * 32;
  const int rem = rem0 + lane;
  const bool valid = rem < a.hw;
  const int w = a.src[0].width[lvl];
  const int pitch = a.src[0].pitch[lvl];
  const float c = valid ? __ldg(a.coords + static_cast<long long>(b) * a.hw + rem) : 0.0f;
  const float centre = __fmul_rn(c, 1.0f / static_cast<float>(1 << lvl));
  const LevelScale sc = level_scale(w, lvl, centre);
  const int s = make_tap(0, R, centre, sc).i0 & ~3;
  const int hi = make_tap(TAPS - 1, R, centre, sc).i1;
  const int q = lane & 3;
  const float* __restrict__ rows = a.src[0].ptr[lvl] + (static_cast<long long>(b) * a.hw + rem0) * pitch;
  float4 v[WINQ];
#pragma unroll
  for (int j = 0; j < WINQ; ++j) {
    const int p = j * 8 + (lane >> 2);
    const int sp = __shfl_sync(FULL, s, p);
    const int hp = __shfl_sync(FULL, hi, p);
    const int cq = sp + 4 * q;
    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((rem0 + p < a.hw) && cq <= hp && cq < w) {
      const float* src = rows + static_cast<long long>(p) * pitch + cq;
      if (a.vec) {
        v[j] = ldg_f4(src);
      } else {
        const int left = w - cq;
        v[j].x = __ldg(src);
        if (left > 1) v[j].y = __ldg(src + 1);
        if (left > 2) v[j].z = __ldg(src + 2);
        if (left > 3) v[j].w = __ldg(src + 3);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < WINQ; ++j) {
    float* dst = win + (j * 8 + (lane >> 2)) * STRIDE + 4 * q;
    dst[0] = v[j].x;
    dst[1] = v[j].y;
    dst[2] = v[j].z;
    dst[3] = v[j].w;
  }
  __syncwarp();
  {
    const float* mine = win + lane * STRIDE - s;
#pragma unroll
    for (int k = 0; k < TAPS; ++k) {
      const Tap tp = make_tap(k, R, centre, sc);
      // coef * val0 + (1 - coef) * val1, each operation rounded (utils.py:27); invalid pixels read zeros
      corr[(lvl * TAPS + k) * 32 + lane] = __fadd_rn(__fmul_rn(tp.coef, mine[tp.i0]), __fmul_rn(tp.one_minus, mine[tp.i1]));
    }
  }
  __syncthreads();  // also orders the weight staging before the first use

  // 1x1 convolution, register tile 4 output channels x 4 pixels per thread: lane = (pixel quad pq = lane & 7,
  // channel quad cq = lane >> 3), warp `lvl` owns channel quads 4*lvl + cq of every 64-channel pass.  Per k:
  // one 16-byte read of W[k][4 channels] (broadcast inside the 8 lanes of a channel quad), one of
  // corr[k][4 pixels], 16 independent FFMA.  Stores: per channel eight lanes write 128 contiguous bytes.
  constexpr int KMAX = 4 * TAPS;  // this kernel is launched with L == 4
  const int pq = lane & 7, cq = lane >> 3;
  const bool vec_out = (a.hw & 3) == 0;
  for (int pass = 0; pass < c_out; pass += 64) {
    const int co0 = pass + 16 * lvl + 4 * cq;   // first of my four output channels
    float acc[4][4];                             // [channel][pixel]
    {
      const float4 bv = co0 < c_out ? *reinterpret_cast<const float4*>(bsm + co0) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = bb[i];
    }
    if (co0 < c_out) {
#pragma unroll 12
      for (int k = 0; k < KMAX; ++k) {
        const float4 wv = *reinterpret_cast<const float4*>(wsm + k * c_pad + co0);
        const float4 xv = *reinterpret_cast<const float4*>(corr + k * 32 + 4 * pq);
        const float ww[4] = {wv.x, wv.y, wv.z, wv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ww[i], xx[j], acc[i][j]);
      }
      float* op = a.out + (static_cast<long long>(b) * c_out + co0) * a.hw + rem0 + 4 * pq;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (co0 + i >= c_out) break;
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = relu ? fmaxf(acc[i][j], 0.f) : acc[i][j];
        float* o = op + static_cast<long long>(i) * a.hw;
        if (vec_out && rem0 + 4 * pq + 3 < a.hw) {
          *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (rem0 + 4 * pq + j < a.hw) o[j] = r[j];
        }
      }
    }
  }
  __syncthreads();  // corr[] is rewritten by the next group
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core variant of the fused lookup + 1x1 convolution (operands rounded to nearest TF32, fp32
// accumulation): the same precision class cuDNN gives this layer when TF32 convolutions are allowed.
// Each warp keeps its share of the weight matrix -- 64 output channels x 40 (36 padded) inputs -- as
// mma.sync m16n8k8 A-fragments in 80 REGISTERS for the whole (persistent) kernel, so the per-group work is
// the lookup itself, 80 MMAs per warp fed from the 36 x 32 tile the lookup left in shared memory, and the
// output stores.  The GEMM all but vanishes; the kernel costs a lookup plus a 61 MB write.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  // not volatile: the scheduler may interleave independent accumulation chains
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int TAPS>
__global__ void __launch_bounds__(128)
corr1d_lookup_conv1x1_tc_kernel(const __grid_constant__ LookupArgs a, const float* __restrict__ weight,
                                const float* __restrict__ bias, int c_out, int relu, int out_nhwc, long long n_groups) {
  constexpr int WINQ = 4, COLS = 4 * WINQ, STRIDE = COLS + 4;  // 16-byte aligned rows for the 16-byte cp.async copies
  constexpr int R = (TAPS - 1) / 2;
  constexpr int K = 4 * TAPS;        // 36 lookup channels (4 levels)
  constexpr int KSTEPS = (K + 7) / 8;  // 5 k-steps of 8, the last one half empty
  constexpr int CS = 40;             // corr tile: [32 px][CS] floats -> B-fragment reads hit 32 distinct banks
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float smem[];
  float* corr = smem;                     // [32][CS]
  float* wins = corr + 32 * CS;           // [2 buffers][4 levels][32][STRIDE]
  const int lane = threadIdx.x, lvl = threadIdx.y;
  const int gid = lane >> 2, tig = lane & 3;

  // A fragments: warp `lvl` owns output channels [64*lvl, 64*lvl + 64): 4 m-tiles x 5 k-steps x 4 registers.
  // weight is k-major (K, c_out).  a0:(row gid, col tig) a1:(gid+8, tig) a2:(gid, tig+4) a3:(gid+8, tig+4)
  uint32_t af[4][KSTEPS][4];
  float bv[4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int r0 = 64 * lvl + 16 * mt + gid, r1 = r0 + 8;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const int k0 = 8 * ks + tig, k1 = k0 + 4;
      // raw loads first (all 80 in flight together), rounding to TF32 in a second sweep below
      af[mt][ks][0] = (r0 < c_out && k0 < K) ? __float_as_uint(__ldg(weight + static_cast<long long>(k0) * c_out + r0)) : 0u;
      af[mt][ks][1] = (r1 < c_out && k0 < K) ? __float_as_uint(__ldg(weight + static_cast<long long>(k0) * c_out + r1)) : 0u;
      af[mt][ks][2] = (r0 < c_out && k1 < K) ? __float_as_uint(__ldg(weight + static_cast<long long>(k1) * c_out + r0)) : 0u;
      af[mt][ks][3] = (r1 < c_out && k1 < K) ? __float_as_uint(__ldg(weight + static_cast<long long>(k1) * c_out + r1)) : 0u;
    }
    bv[mt][0] = (bias && r0 < c_out) ? __ldg(bias + r0) : 0.f;
    bv[mt][1] = (bias && r1 < c_out) ? __ldg(bias + r1) : 0.f;
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
      for (int e = 0; e < 4; ++e) af[mt][ks][e] = to_tf32(__uint_as_float(af[mt][ks][e]));
  // zero the padding columns of the corr tile once (k = 36..39 feed the last k-step)
  for (int i = lvl * 32 + lane; i < 32 * (CS - K); i += 128) corr[(i / (CS - K)) * CS + K + i % (CS - K)] = 0.f;

  const int groups_per_image = (a.hw + 31) / 32;
  const int w = a.src[0].width[lvl];
  const int pitch = a.src[0].pitch[lvl];
  const LevelScale sc = level_scale(w, lvl, 0.f);
  const int q = lane & 3;

  // Software pipeline over the block's pixel groups: coordinates are fetched two groups ahead, windows one
  // group ahead (cp.async into the other half of the double-buffered window tile, issued right before the MMA
  // phase of the current group), so neither DRAM latency sits on the per-group critical path.
  auto group_coord = [&](long long grp, int& b, int& rem0) -> float {
    if (grp >= n_groups) { b = 0; rem0 = 0; return 0.0f; }
    b = static_cast<int>(grp / groups_per_image);
    rem0 = static_cast<int>(grp - static_cast<long long>(b) * groups_per_image) * 32;
    const int rem = rem0 + lane;
    return rem < a.hw ? __ldg(a.coords + static_cast<long long>(b) * a.hw + rem) : 0.0f;
  };
  auto issue_windows = [&](float c, int b, int rem0, float* win) {
    const float centre = __fmul_rn(c, 1.0f / static_cast<float>(1 << lvl));
    const int s = make_tap(0, R, centre, sc).i0 & ~3;
    const int hi = make_tap(TAPS - 1, R, centre, sc).i1;
    const float* __restrict__ rows = a.src[0].ptr[lvl] + (static_cast<long long>(b) * a.hw + rem0) * pitch;
#pragma unroll
    for (int j = 0; j < WINQ; ++j) {
      const int p = j * 8 + (lane >> 2);
      const int sp = __shfl_sync(FULL, s, p);
      const int hp = __shfl_sync(FULL, hi, p);
      const int cq = sp + 4 * q;
      if ((rem0 + p < a.hw) && cq <= hp && cq < w) {
        const float* src = rows + static_cast<long long>(p) * pitch + cq;
        const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(win + p * STRIDE + 4 * q));
        if (a.vec) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        } else {
          const int left = w - cq;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (e < left) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * e), "l"(src + e) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const long long stride = gridDim.x;
  long long grp = blockIdx.x;
  int b, rem0, b1, rem01, b2, rem02;
  float c = group_coord(grp, b, rem0);
  float c1 = group_coord(grp + stride, b1, rem01);
  int buf = 0;
  if (grp < n_groups) issue_windows(c, b, rem0, wins + (buf * 4 + lvl) * (32 * STRIDE));
  for (; grp < n_groups; grp += stride) {
    const float c2 = group_coord(grp + 2 * stride, b2, rem02);   // in flight during this whole iteration
    const float* win = wins + (buf * 4 + lvl) * (32 * STRIDE);
    const float centre = __fmul_rn(c, 1.0f / static_cast<float>(1 << lvl));
    const int s = make_tap(0, R, centre, sc).i0 & ~3;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    {
      const float* mine = win + lane * STRIDE - s;
#pragma unroll
      for (int k = 0; k < TAPS; ++k) {
        const Tap tp = make_tap(k, R, centre, sc);
        // coef * val0 + (1 - coef) * val1, each operation rounded (utils.py:27); then RN to TF32 for the MMA
        const float val = __fadd_rn(__fmul_rn(tp.coef, mine[tp.i0]), __fmul_rn(tp.one_minus, mine[tp.i1]));
        corr[lane * CS + lvl * TAPS + k] = __uint_as_float(to_tf32(val));
      }
    }
    __syncthreads();
    if (grp + stride < n_groups) issue_windows(c1, b1, rem01, wins + ((buf ^ 1) * 4 + lvl) * (32 * STRIDE));

    // D[co][px] = sum_k W[co][k] * corr[px][k]: 4 m-tiles x 4 n-tiles (8 pixels each) per warp
    float* out_img = a.out + static_cast<long long>(b) * c_out * a.hw + rem0;
    const bool full_tile = rem0 + 32 <= a.hw && 64 * lvl + 64 <= c_out && (c_out & 1) == 0 && (a.hw & 1) == 0;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t bf[KSTEPS][2];   // b0:(k = tig, n = gid)  b1:(k = tig + 4, n = gid)
      const float* cp = corr + (8 * nt + gid) * CS + tig;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        bf[ks][0] = __float_as_uint(cp[8 * ks]);
        bf[ks][1] = __float_as_uint(cp[8 * ks + 4]);
      }
      const int px = 8 * nt + 2 * tig;   // my two output pixels (c0/c1 and c2/c3 columns)
      // the four m-tiles are independent accumulation chains: k-step outermost, so consecutive MMAs never depend
      float dd[4][4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        dd[mt][0] = dd[mt][1] = bv[mt][0];
        dd[mt][2] = dd[mt][3] = bv[mt][1];
      }
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) mma_tf32_16x8x8(dd[mt], af[mt][ks], bf[ks][0], bf[ks][1]);
      if (full_tile) {
        // whole group inside the image and all 64 channels of this warp exist: one base pointer per n-tile, constant
        // offsets per m-tile, no predicates
        const long long pix0 = static_cast<long long>(b) * a.hw + rem0 + px;
        if (out_nhwc == 2) {
          const bool odd = gid & 1;
          uint16_t* ob = reinterpret_cast<uint16_t*>(a.out) + (pix0 + (odd ? 1 : 0)) * c_out + 64 * lvl + (gid & ~1);
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            float (&d)[4] = dd[mt];
            if (relu) {
#pragma unroll
              for (int i = 0; i < 4; ++i) d[i] = fmaxf(d[i], 0.f);
            }
            const float s0 = __shfl_xor_sync(0xffffffffu, odd ? d[0] : d[1], 4);
            const float s1 = __shfl_xor_sync(0xffffffffu, odd ? d[2] : d[3], 4);
            *reinterpret_cast<uint32_t*>(ob + 16 * mt) = odd ? pack_h2(s0, d[1]) : pack_h2(d[0], s0);
            *reinterpret_cast<uint32_t*>(ob + 16 * mt + 8) = odd ? pack_h2(s1, d[3]) : pack_h2(d[2], s1);
          }
        } else if (out_nhwc == 1) {
          float* ob = a.out + pix0 * c_out + 64 * lvl + gid;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            float (&d)[4] = dd[mt];
            if (relu) {
#pragma unroll
              for (int i = 0; i < 4; ++i) d[i] = fmaxf(d[i], 0.f);
            }
            ob[16 * mt] = d[0];
            ob[16 * mt + 8] = d[2];
            ob[c_out + 16 * mt] = d[1];
            ob[c_out + 16 * mt + 8] = d[3];
          }
        } else {
          float* ob = out_img + static_cast<long long>(64 * lvl + gid) * a.hw + px;
          const long long hw8 = 8LL * a.hw;
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            float (&d)[4] = dd[mt];
            if (relu) {
#pragma unroll
              for (int i = 0; i < 4; ++i) d[i] = fmaxf(d[i], 0.f);
            }
            *reinterpret_cast<float2*>(ob + (2 * mt) * hw8) = make_float2(d[0], d[1]);
            *reinterpret_cast<float2*>(ob + (2 * mt + 1) * hw8) = make_float2(d[2], d[3]);
          }
        }
        continue;
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        float (&d)[4] = dd[mt];
        if (relu) {
#pragma unroll
          for (int i = 0; i < 4; ++i) d[i] = fmaxf(d[i], 0.f);
        }
        const int r0 = 64 * lvl + 16 * mt + gid, r1 = r0 + 8;
        if (out_nhwc) {
          // channels-last output (B, H*W, c_out): eight lanes (gid) write 32 contiguous bytes of one pixel
          if (out_nhwc == 2) {
            // the same, rounded to IEEE fp16 (the consumer runs as an fp16 convolution).  Lanes gid and gid^1 hold
            // adjacent channels of the same two pixels: one exchange each way and the even lane owns channel pairs
            // (r, r+1) of pixel px, the odd lane those of pixel px+1 -> 4-byte stores, 16 contiguous bytes per pixel
            const bool odd = gid & 1;
            const float s0 = __shfl_xor_sync(0xffffffffu, odd ? d[0] : d[1], 4);
            const float s1 = __shfl_xor_sync(0xffffffffu, odd ? d[2] : d[3], 4);
            const uint32_t lo = odd ? pack_h2(s0, d[1]) : pack_h2(d[0], s0);   // channels (r0e, r0e + 1)
            const uint32_t hi = odd ? pack_h2(s1, d[3]) : pack_h2(d[2], s1);   // channels (r0e + 8, r0e + 9)
            const int re = r0 - (odd ? 1 : 0);                                  // even channel of the pair
            const long long pix = rem0 + px + (odd ? 1 : 0);
            if (pix < a.hw) {
              uint16_t* o16 = reinterpret_cast<uint16_t*>(a.out) + (static_cast<long long>(b) * a.hw + pix) * c_out;
              if ((c_out & 1) == 0) {
                if (re < c_out) *reinterpret_cast<uint32_t*>(o16 + re) = lo;
                if (re + 8 < c_out) *reinterpret_cast<uint32_t*>(o16 + re + 8) = hi;
              } else {
                if (re < c_out) o16[re] = static_cast<uint16_t>(lo);
                if (re + 1 < c_out) o16[re + 1] = static_cast<uint16_t>(lo >> 16);
                if (re + 8 < c_out) o16[re + 8] = static_cast<uint16_t>(hi);
                if (re + 9 < c_out) o16[re + 9] = static_cast<uint16_t>(hi >> 16);
              }
            }
            continue;
          }
          float* o = a.out + (static_cast<long long>(b) * a.hw + rem0 + px) * c_out;
          if (rem0 + px < a.hw) {
            if (r0 < c_out) o[r0] = d[0];
            if (r1 < c_out) o[r1] = d[2];
          }
          if (rem0 + px + 1 < a.hw) {
            if (r0 < c_out) o[c_out + r0] = d[1];
            if (r1 < c_out) o[c_out + r1] = d[3];
          }
          continue;
        }
        float* o0 = out_img + static_cast<long long>(r0) * a.hw + px;
        float* o1 = out_img + static_cast<long long>(r1) * a.hw + px;
        if (rem0 + px + 1 < a.hw && (a.hw & 1) == 0) {
          if (r0 < c_out) *reinterpret_cast<float2*>(o0) = make_float2(d[0], d[1]);
          if (r1 < c_out) *reinterpret_cast<float2*>(o1) = make_float2(d[2], d[3]);
        } else {
          if (rem0 + px < a.hw) {
            if (r0 < c_out) o0[0] = d[0];
            if (r1 < c_out) o1[0] = d[2];
          }
          if (rem0 + px + 1 < a.hw) {
            if (r0 < c_out) o0[1] = d[1];
            if (r1 < c_out) o1[1] = d[3];
          }
        }
      }
    }
    __syncthreads();  // corr[] is rewritten by the next group
    c = c1; b = b1; rem0 = rem01;
    c1 = c2; b1 = b2; rem01 = rem02;
    buf ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------
// IGEV dual lookup over the interleaved pyramids ([b][h][w1][d][g], igev.cu): the eight windows of a
// pixel are one contiguous run of (hi - lo + 1) * 32 bytes <= 352 bytes, so every fetched sector is used.
// Warp = 32 consecutive pixels x one level; for each of the two sources the warp copies the 32 runs into a
// padded shared-memory tile with 16-byte loads (8 lanes per pixel, 4 pixels per instruction), then every
// lane interpolates its own pixel: per tap two 32-byte reads give the 8 groups of idx0 and idx1.
// Output layout and arithmetic are those of the reference (channel = l*144 + src*72 + g*9 + k).
// ------------------------------------------------------------------------------------------------
struct GevLookupArgs {
  const float* src[2][4];  // [source][level]
  int width[4];            // D >> l
  const float* coords;
  float* out;
  int hw;
  int num_levels;
};

constexpr int GEV_G = 8;
constexpr int GEV_T = 9;           // radius 4
constexpr int GEV_RUN = 11 * GEV_G;   // floats in the longest run
constexpr int GEV_STRIDE = GEV_RUN + 4;  // 92 floats: 16-byte aligned rows, 28-bank skew -> conflict-free LDS.128

__global__ void __launch_bounds__(128)
gev_lookup_kernel(const __grid_constant__ GevLookupArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float gsm[];
  const int lane = threadIdx.x;
  const int lvl = threadIdx.y;
  float* win = gsm + lvl * (32 * GEV_STRIDE);
  const int b = blockIdx.y;
  const int rem0 = blockIdx.x * 32;
  const int rem = rem0 + lane;
  const bool valid = rem < a.hw;
  const int w = a.width[lvl];

  const float c = valid ? __ldg(a.coords + static_cast<long long>(b) * a.hw + rem) : 0.0f;
  const float centre = __fmul_rn(c, 1.0f / static_cast<float>(1 << lvl));
  const LevelScale sc = level_scale(w, lvl, centre);
  int o0[GEV_T], o1[GEV_T];
  float cf[GEV_T], omc[GEV_T];
  int lo, hi;
  {
    Tap tp[GEV_T];
#pragma unroll
    for (int k = 0; k < GEV_T; ++k) tp[k] = make_tap(k, 4, centre, sc);
    lo = tp[0].i0;
    hi = tp[GEV_T - 1].i1;
    const int base = lane * GEV_STRIDE - lo * GEV_G;
#pragma unroll
    for (int k = 0; k < GEV_T; ++k) {
      o0[k] = base + tp[k].i0 * GEV_G;
      o1[k] = base + tp[k].i1 * GEV_G;
      cf[k] = tp[k].coef;
      omc[k] = tp[k].one_minus;
    }
  }
  const int nq = valid ? 2 * (hi - lo + 1) : 0;  // 16-byte quads in my run (<= 22)

  // loader role: round r serves pixels 4r .. 4r+3, eight lanes each, three quads per lane
  const int sub = lane & 7;
  const long long rowlen = static_cast<long long>(w) * GEV_G;
  const long long pix_base = static_cast<long long>(b) * a.hw + rem0;
  const long long c_total = static_cast<long long>(a.num_levels) * 2 * GEV_G * GEV_T;
  float* const out_px = a.out + static_cast<long long>(b) * c_total * a.hw + rem;

  for (int s = 0; s < 2; ++s) {
    const float* __restrict__ rows = a.src[s][lvl] + pix_base * rowlen;
    // 16-byte cp.async copies straight into the tile: all 24 of a lane are in flight at once and cost no
    // registers (through registers ptxas interleaves "load round r / store round r-1" and keeps ~3 in flight)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int p = 4 * r + (lane >> 3);
      const int lo_p = __shfl_sync(FULL, lo, p);
      const int nq_p = __shfl_sync(FULL, nq, p);
      const float* src = rows + static_cast<long long>(p) * rowlen + lo_p * GEV_G;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(win + p * GEV_STRIDE));
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int q = sub + 8 * j;
        if (q < nq_p)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * q), "l"(src + 4 * q) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (valid) {
      float* op = out_px + (static_cast<long long>(lvl) * 2 * GEV_G * GEV_T + s * GEV_G * GEV_T) * a.hw;
#pragma unroll
      for (int k = 0; k < GEV_T; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(win + o0[k]);
        const float4 a1 = *reinterpret_cast<const float4*>(win + o0[k] + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(win + o1[k]);
        const float4 b1 = *reinterpret_cast<const float4*>(win + o1[k] + 4);
        const float v0[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float v1[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int g = 0; g < GEV_G; ++g)
          // coef * val0 + (1 - coef) * val1, each operation rounded (utils.py:27)
          op[static_cast<long long>(g * GEV_T + k) * a.hw] = __fadd_rn(__fmul_rn(cf[k], v0[g]), __fmul_rn(omc[k], v1[g]));
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of the pyramid lookup with respect to the pyramid (training; the coordinates are detached by the
// reference before every lookup, raft_stereo/model.py:131).  out = coef * row[i0] + (1 - coef) * row[i1] per tap,
// so d_row[i0] += coef * g and d_row[i1] += (1 - coef) * g.  A pyramid row belongs to exactly one pixel, so the
// thread that owns (pixel, level) sums its <= 2r+3 window entries in registers and writes them with plain stores:
// no atomics, deterministic.  The caller hands in zero-initialised gradient levels (entries outside the windows
// stay zero).
// ------------------------------------------------------------------------------------------------
struct LookupBwdArgs {
  float* d_level[NND_MAX_LEVELS];
  int width[NND_MAX_LEVELS];
  int pitch[NND_MAX_LEVELS];
  const float* coords;
  const float* grad_out;
  long long n_pix;   // B * H * W1
  int hw;
  int num_levels;
  int radius;
};

__global__ void __launch_bounds__(256)
corr1d_lookup_backward_kernel(const __grid_constant__ LookupBwdArgs a) {
  constexpr int MAXWIN = 32;  // 2r + 3 <= 29 for r <= 13
  const int T = 2 * a.radius + 1;
  const long long total = a.n_pix * a.num_levels;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int lvl = static_cast<int>(i / a.n_pix);
    const long long pix = i - static_cast<long long>(lvl) * a.n_pix;
    const long long b = pix / a.hw;
    const int rem = static_cast<int>(pix - b * a.hw);
    const int w = a.width[lvl];
    const float centre = __fmul_rn(__ldg(a.coords + pix), 1.0f / static_cast<float>(1 << lvl));
    const LevelScale sc = level_scale(w, lvl, centre);
    const int lo = make_tap(0, a.radius, centre, sc).i0;
    float acc[MAXWIN];
#pragma unroll
    for (int j = 0; j < MAXWIN; ++j) acc[j] = 0.f;
    const float* g = a.grad_out + (b * a.num_levels + lvl) * T * a.hw + rem;
    for (int k = 0; k < T; ++k) {
      const Tap tp = make_tap(k, a.radius, centre, sc);
      const float gk = __ldg(g + static_cast<long long>(k) * a.hw);
      const int j0 = tp.i0 - lo, j1 = tp.i1 - lo;
#pragma unroll
      for (int j = 0; j < MAXWIN; ++j) {   // register array: select instead of dynamic indexing
        if (j == j0) acc[j] = fmaf(tp.coef, gk, acc[j]);
        if (j == j1) acc[j] = fmaf(tp.one_minus, gk, acc[j]);
      }
    }
    float* row = a.d_level[lvl] + pix * a.pitch[lvl] + lo;
    const int n = min(w - lo, MAXWIN);
#pragma unroll
    for (int j = 0; j < MAXWIN; ++j)
      if (j < n && j <= 2 * a.radius + 2) row[j] = acc[j];
  }
}

// backward of avg_pool1d(., 2): d_src[r][2j] += d_dst[r][j] / 2, d_src[r][2j+1] += d_dst[r][j] / 2
__global__ void avgpool_pairs_backward_kernel(const float* __restrict__ d_dst, int dst_width, int dst_pitch,
                                              float* __restrict__ d_src, int src_pitch, long long rows) {
  const long long total = rows * dst_width;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dst_width;
    const int j = static_cast<int>(i - r * dst_width);
    const float h = 0.5f * __ldg(d_dst + r * dst_pitch + j);
    float* s = d_src + r * src_pitch + 2 * j;
    s[0] += h;
    s[1] += h;
  }
}

__global__ void lookup_indices_kernel(const float* __restrict__ coords, long long n_pix, int num_levels, int radius,
                                      int w0, int w1, int w2, int w3, int w4, int w5, int w6, int w7,
                                      int32_t* __restrict__ idx0, int32_t* __restrict__ idx1) {
  const int T = 2 * radius + 1;
  const long long total = n_pix * T * num_levels;
  const int widths[NND_MAX_LEVELS] = {w0, w1, w2, w3, w4, w5, w6, w7};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % T);
    const long long pl = i / T;
    const long long pix = pl % n_pix;
    const int lvl = static_cast<int>(pl / n_pix);
    const float centre = __fmul_rn(__ldg(coords + pix), 1.0f / static_cast<float>(1 << lvl));
    const Tap tp = make_tap<true>(k, radius, centre, level_scale(widths[lvl], lvl, centre));
    idx0[i] = tp.i0;
    idx1[i] = tp.i1;
  }
}

static nnd_status launch_lookup(const float* const* level_a, const float* const* level_b, const int* width,
                                const int* pitch, const float* coords, int B, int G, int H, int W1, int num_levels,
                                int radius, int n_src, int mode, float* out, cudaStream_t stream) {
  NND_REQUIRE(level_a && width && pitch && coords && out, "lookup: null pointer argument");
  NND_REQUIRE(B > 0 && G > 0 && H > 0 && W1 > 0, "lookup: B, G, H, W1 must be positive (got %d %d %d %d)", B, G, H, W1);
  NND_REQUIRE(B <= 65535, "lookup: batch %d exceeds the grid limit (65535)", B);
  NND_REQUIRE(static_cast<long long>(H) * W1 < (1LL << 30), "lookup: H*W1 too large");
  NND_REQUIRE(num_levels >= 1 && num_levels <= NND_MAX_LEVELS, "lookup: num_levels %d outside [1, %d]", num_levels,
              NND_MAX_LEVELS);
  NND_REQUIRE(radius >= 0 && radius <= 13, "lookup: radius %d outside [0, 13]", radius);
  NND_REQUIRE(n_src == 1 || level_b, "lookup: second pyramid missing");

  LookupArgs a;
  memset(&a, 0, sizeof(a));
  bool vec = true;
  for (int l = 0; l < num_levels; ++l) {
    // linear_sampler divides by (w2 - 1): a 1-wide level is a division by zero in the reference
    NND_REQUIRE(width[l] >= 2, "lookup: level %d has width %d; linear_sampler needs width >= 2", l, width[l]);
    NND_REQUIRE(width[l] <= (1 << 22), "lookup: level %d width %d too large", l, width[l]);
    NND_REQUIRE(pitch[l] >= width[l], "lookup: level %d pitch %d < width %d", l, pitch[l], width[l]);
    NND_REQUIRE(level_a[l], "lookup: level %d pointer is null", l);
    a.src[0].ptr[l] = level_a[l];
    a.src[0].width[l] = width[l];
    a.src[0].pitch[l] = pitch[l];
    vec = vec && (pitch[l] % 4 == 0) && aligned16(level_a[l]);
    if (n_src == 2) {
      NND_REQUIRE(level_b[l], "lookup: geometry level %d pointer is null", l);
      a.src[1].ptr[l] = level_b[l];
      a.src[1].width[l] = width[l];
      a.src[1].pitch[l] = pitch[l];
      vec = vec && aligned16(level_b[l]);
    }
  }
  a.coords = coords;
  a.out = out;
  a.hw = H * W1;
  a.G = G;
  a.n_src = n_src;
  a.num_levels = num_levels;
  a.radius = radius;
  a.mode = mode;
  a.vec = vec ? 1 : 0;
  const int n_planes = n_src * G;
  a.planes_per_block = n_planes >= 8 ? 4 : n_planes;

  if (n_src == 1 && G == 1 && mode == 0 && radius == 4) {
    // RAFT-Stereo: register-lean variant, whole grid resident in one wave (see corr1d_lookup_lean_kernel)
    dim3 grid((a.hw + 31) / 32, B);
    dim3 block(32, num_levels);
    const size_t smem = static_cast<size_t>(num_levels) * 32 * 17 * sizeof(float);
    NND_REQUIRE(width[0] <= 16384, "lookup: level-0 width %d exceeds 16384", width[0]);   // conservative window start
    corr1d_lookup_lean_kernel<9><<<grid, block, smem, stream>>>(a, 0x4b000000u << 2);
    return check_launch("corr1d_lookup_lean_kernel");
  }
  dim3 grid((a.hw + 31) / 32, (n_planes + a.planes_per_block - 1) / a.planes_per_block, B);
  dim3 block(32, num_levels);
  if (radius <= 5) {
    const size_t smem = static_cast<size_t>(num_levels) * 17 * 32 * sizeof(float);
    if (radius == 4 && mode == 0)
      pyramid_lookup_kernel<4, 9><<<grid, block, smem, stream>>>(a);
    else
      pyramid_lookup_kernel<4, 0><<<grid, block, smem, stream>>>(a);
  } else {
    const size_t smem = static_cast<size_t>(num_levels) * 33 * 32 * sizeof(float);
    pyramid_lookup_kernel<8, 0><<<grid, block, smem, stream>>>(a);
  }
  return check_launch("pyramid_lookup_kernel");
}

}  // namespace nnd

namespace nnd {
nnd_status launch_lookup_conv1x1_ws(const LookupArgs& a, const float* weight, const float* bias, int relu, int out_f16,
                                    long long total_px, int skew_w1, cudaStream_t stream);
nnd_status launch_lookup_skewed(const LookupArgs& a, int B, int H, int W1, cudaStream_t stream);
nnd_status launch_corr1d_skew(const ConstPyramid& src, int num_levels, int B, int H, int W1, float* const* dst, int P1,
                              cudaStream_t stream);
}

extern "C" {

nnd_status nnd_corr1d_lookup(const float* const* level, const int* width, const int* pitch, const float* coords,
                             int B, int H, int W1, int num_levels, int radius, float* out, nnd_stream_t stream) {
  return nnd::launch_lookup(level, nullptr, width, pitch, coords, B, 1, H, W1, num_levels, radius, 1, 0, out,
                            reinterpret_cast<cudaStream_t>(stream));
}

nnd_status nnd_group_lookup(const float* const* level_a, const float* const* level_b, const int* width,
                            const int* pitch, const float* coords, int B, int G, int H, int W1, int num_levels,
                            int radius, int mode, float* out, nnd_stream_t stream) {
  if (mode != 0 && mode != 1) {
    nnd::set_error("group_lookup: mode %d is not 0 (IGEV dual) or 1 (GroupCorrBlock1D)", mode);
    return NND_ERR_INVALID_ARGUMENT;
  }
  return nnd::launch_lookup(level_a, mode == 0 ? level_b : nullptr, width, pitch, coords, B, G, H, W1, num_levels,
                            radius, mode == 0 ? 2 : 1, mode, out, reinterpret_cast<cudaStream_t>(stream));
}

nnd_status nnd_corr1d_lookup_indices(const int* width, const float* coords, int B, int H, int W1, int num_levels,
                                     int radius, int32_t* idx0, int32_t* idx1, nnd_stream_t stream) {
  NND_REQUIRE(width && coords && idx0 && idx1, "lookup_indices: null pointer argument");
  NND_REQUIRE(B > 0 && H > 0 && W1 > 0, "lookup_indices: B, H, W1 must be positive");
  NND_REQUIRE(num_levels >= 1 && num_levels <= NND_MAX_LEVELS, "lookup_indices: num_levels %d outside [1, %d]",
              num_levels, NND_MAX_LEVELS);
  NND_REQUIRE(radius >= 0 && radius <= 64, "lookup_indices: radius %d outside [0, 64]", radius);
  int w[NND_MAX_LEVELS] = {2, 2, 2, 2, 2, 2, 2, 2};
  for (int l = 0; l < num_levels; ++l) {
    NND_REQUIRE(width[l] >= 2 && width[l] <= (1 << 22),
                "lookup_indices: level %d has width %d; linear_sampler needs width >= 2", l, width[l]);
    w[l] = width[l];
  }
  const long long n_pix = static_cast<long long>(B) * H * W1;
  const long long total = n_pix * (2 * radius + 1) * num_levels;
  const int threads = 256;
  const long long want = (total + threads - 1) / threads;
  const int blocks = static_cast<int>(want < 148LL * 32 ? want : 148LL * 32);
  nnd::lookup_indices_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      coords, n_pix, num_levels, radius, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], idx0, idx1);
  return nnd::check_launch("lookup_indices_kernel");
}

nnd_status nnd_gev_lookup(const float* const* level_feat, const float* const* level_geo, const float* coords, int B,
                          int G, int D, int H, int W1, int num_levels, int radius, float* out, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(level_feat && level_geo && coords && out, "gev_lookup: null pointer argument");
  NND_REQUIRE(B > 0 && H > 0 && W1 > 0, "gev_lookup: B, H, W1 must be positive");
  NND_REQUIRE(B <= 65535, "gev_lookup: batch %d exceeds the grid limit (65535)", B);
  NND_REQUIRE(G == GEV_G && radius == 4, "gev_lookup: built for 8 groups and radius 4 (got G=%d, radius=%d)", G, radius);
  NND_REQUIRE(num_levels >= 1 && num_levels <= 4, "gev_lookup: num_levels %d outside [1, 4]", num_levels);
  NND_REQUIRE(D % 8 == 0, "gev_lookup: D = %d must be a multiple of 8", D);
  NND_REQUIRE(static_cast<long long>(H) * W1 < (1LL << 30), "gev_lookup: H*W1 too large");
  GevLookupArgs a;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < num_levels; ++l) {
    // linear_sampler divides by (w2 - 1): a 1-wide level is a division by zero in the reference
    NND_REQUIRE((D >> l) >= 2, "gev_lookup: level %d has width %d; linear_sampler needs width >= 2", l, D >> l);
    NND_REQUIRE(level_feat[l] && level_geo[l] && aligned16(level_feat[l]) && aligned16(level_geo[l]),
                "gev_lookup: level %d pointer is null or unaligned", l);
    a.src[0][l] = level_feat[l];
    a.src[1][l] = level_geo[l];
    a.width[l] = D >> l;
  }
  a.coords = coords;
  a.out = out;
  a.hw = H * W1;
  a.num_levels = num_levels;
  dim3 grid((a.hw + 31) / 32, B);
  dim3 block(32, num_levels);
  const size_t smem = static_cast<size_t>(num_levels) * 32 * GEV_STRIDE * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(gev_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  gev_lookup_kernel<<<grid, block, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("gev_lookup_kernel");
}

nnd_status nnd_corr1d_lookup_conv1x1(const float* const* level, const int* width, const int* pitch, const float* coords,
                                     int B, int H, int W1, int num_levels, int radius, const float* weight,
                                     const float* bias, int c_out, int relu, int precision, int out_layout,
                                     void* out, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(level && width && pitch && coords && weight && out, "lookup_conv1x1: null pointer argument");
  NND_REQUIRE(B > 0 && H > 0 && W1 > 0 && c_out > 0, "lookup_conv1x1: B, H, W1, c_out must be positive");
  NND_REQUIRE(B <= 65535, "lookup_conv1x1: batch %d exceeds the grid limit (65535)", B);
  NND_REQUIRE(radius == 4, "lookup_conv1x1: built for radius 4 (got %d)", radius);
  NND_REQUIRE(num_levels == 4, "lookup_conv1x1: built for the 4-level pyramid (got %d levels)", num_levels);
  NND_REQUIRE(static_cast<long long>(H) * W1 < (1LL << 30), "lookup_conv1x1: H*W1 too large");
  LookupArgs a;
  memset(&a, 0, sizeof(a));
  bool vec = true;
  for (int l = 0; l < num_levels; ++l) {
    NND_REQUIRE(width[l] >= 2, "lookup_conv1x1: level %d has width %d; linear_sampler needs width >= 2", l, width[l]);
    NND_REQUIRE(pitch[l] >= width[l] && level[l], "lookup_conv1x1: level %d pointer/pitch invalid", l);
    a.src[0].ptr[l] = level[l];
    a.src[0].width[l] = width[l];
    a.src[0].pitch[l] = pitch[l];
    vec = vec && (pitch[l] % 4 == 0) && aligned16(level[l]);
  }
  a.coords = coords;
  a.out = reinterpret_cast<float*>(out);
  a.hw = H * W1;
  a.G = 1;
  a.n_src = 1;
  a.num_levels = num_levels;
  a.radius = radius;
  a.vec = vec ? 1 : 0;
  NND_REQUIRE(precision == NND_PREC_FP32 || precision == NND_PREC_TF32, "lookup_conv1x1: unknown precision %d", precision);
  NND_REQUIRE(out_layout >= 0 && out_layout <= 2,
              "lookup_conv1x1: out_layout %d is not 0 (fp32 NCHW), 1 (fp32 channels-last) or 2 (fp16 channels-last)", out_layout);
  const long long n_groups_all = static_cast<long long>(B) * ((a.hw + 31) / 32);
  if (precision == NND_PREC_TF32 && c_out == 256 && out_layout != 0 && vec) {
    // the shipping shape: warp-specialised tcgen05 kernel (lookup_ws.cu) -- weights in shared memory, accumulators in
    // TMEM, producer / MMA / epilogue warps decoupled by mbarrier pipelines
    return launch_lookup_conv1x1_ws(a, weight, bias, relu ? 1 : 0, out_layout == 2 ? 1 : 0,
                                    static_cast<long long>(B) * a.hw, 0, reinterpret_cast<cudaStream_t>(stream));
  }
  if (precision == NND_PREC_TF32 && c_out <= 256) {
    // tensor-core path: weights live in registers, shared memory holds only the lookup tiles
    const size_t smem_tc = (32 * 40 + static_cast<size_t>(2) * 4 * 32 * 20) * sizeof(float);
    const long long resident = static_cast<long long>(sm_count()) * 3;   // 128 threads x ~150 registers
    dim3 grid_tc(static_cast<unsigned>(n_groups_all < resident ? n_groups_all : resident));
    corr1d_lookup_conv1x1_tc_kernel<9><<<grid_tc, dim3(32, 4), smem_tc, reinterpret_cast<cudaStream_t>(stream)>>>(
        a, weight, bias, c_out, relu ? 1 : 0, out_layout, n_groups_all);
    return check_launch("corr1d_lookup_conv1x1_tc_kernel");
  }
  NND_REQUIRE(out_layout == 0, "lookup_conv1x1: the channels-last outputs are provided by the tensor-core path only");
  const int K = num_levels * 9;
  const size_t c_pad = (static_cast<size_t>(c_out) + 3) & ~static_cast<size_t>(3);
  const size_t smem = (c_pad * K + c_pad + static_cast<size_t>(K) * 32 + static_cast<size_t>(num_levels) * 32 * 17) *
                      sizeof(float);
  NND_REQUIRE(smem <= 200 * 1024, "lookup_conv1x1: c_out = %d needs %zu bytes of shared memory", c_out, smem);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(corr1d_lookup_conv1x1_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const long long n_groups = static_cast<long long>(B) * ((a.hw + 31) / 32);
  const long long resident = static_cast<long long>(sm_count()) * (smem > 100 * 1024 ? 1 : smem > 70 * 1024 ? 2 : smem > 52 * 1024 ? 3 : 4);
  dim3 grid(static_cast<unsigned>(n_groups < resident ? n_groups : resident));
  dim3 block(32, num_levels);
  corr1d_lookup_conv1x1_kernel<9><<<grid, block, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a, weight, bias, c_out,
                                                                                                 relu ? 1 : 0, n_groups);
  return check_launch("corr1d_lookup_conv1x1_kernel");
}

nnd_status nnd_corr1d_lookup_backward(const float* grad_out, const float* coords, const int* width, const int* pitch, int B,
                                      int H, int W1, int num_levels, int radius, float* const* d_level,
                                      nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(grad_out && coords && width && pitch && d_level, "lookup_backward: null pointer argument");
  NND_REQUIRE(B > 0 && H > 0 && W1 > 0, "lookup_backward: B, H, W1 must be positive");
  NND_REQUIRE(num_levels >= 1 && num_levels <= NND_MAX_LEVELS, "lookup_backward: num_levels %d outside [1, %d]", num_levels,
              NND_MAX_LEVELS);
  NND_REQUIRE(radius >= 0 && radius <= 13, "lookup_backward: radius %d outside [0, 13]", radius);
  NND_REQUIRE(static_cast<long long>(H) * W1 < (1LL << 30), "lookup_backward: H*W1 too large");
  LookupBwdArgs a;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < num_levels; ++l) {
    NND_REQUIRE(width[l] >= 2 && pitch[l] >= width[l] && d_level[l], "lookup_backward: level %d invalid", l);
    a.d_level[l] = d_level[l];
    a.width[l] = width[l];
    a.pitch[l] = pitch[l];
  }
  a.coords = coords;
  a.grad_out = grad_out;
  a.hw = H * W1;
  a.n_pix = static_cast<long long>(B) * a.hw;
  a.num_levels = num_levels;
  a.radius = radius;
  const long long total = a.n_pix * num_levels;
  const long long want = (total + 255) / 256, cap = static_cast<long long>(sm_count()) * 8;
  corr1d_lookup_backward_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("corr1d_lookup_backward_kernel");
}

nnd_status nnd_avgpool_pairs_backward(const float* d_dst, int dst_width, int dst_pitch, float* d_src, int src_pitch,
                                      int64_t rows, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(d_dst && d_src, "avgpool_pairs_backward: null pointer");
  NND_REQUIRE(dst_width >= 1 && rows > 0 && dst_pitch >= dst_width && src_pitch >= 2 * dst_width,
              "avgpool_pairs_backward: widths / pitches inconsistent");
  const long long total = rows * dst_width;
  const long long want = (total + 255) / 256, cap = static_cast<long long>(sm_count()) * 16;
  avgpool_pairs_backward_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_dst, dst_width, dst_pitch, d_src, src_pitch, rows);
  return check_launch("avgpool_pairs_backward_kernel");
}

nnd_status nnd_corr1d_skew(const float* const* level, const int* width, const int* pitch, int B, int H, int W1, int num_levels,
                           float* const* skewed, int skew_pitch, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(level && width && pitch && skewed, "corr1d_skew: null pointer argument");
  NND_REQUIRE(B > 0 && H > 0 && W1 > 0, "corr1d_skew: B, H, W1 must be positive");
  NND_REQUIRE(num_levels >= 1 && num_levels <= NND_MAX_LEVELS, "corr1d_skew: num_levels %d outside [1, %d]", num_levels, NND_MAX_LEVELS);
  NND_REQUIRE(skew_pitch >= W1, "corr1d_skew: skew_pitch %d smaller than W1 %d", skew_pitch, W1);
  ConstPyramid src;
  memset(&src, 0, sizeof(src));
  for (int l = 0; l < num_levels; ++l) {
    NND_REQUIRE(level[l] && skewed[l] && width[l] >= 1 && pitch[l] >= width[l], "corr1d_skew: level %d invalid", l);
    src.ptr[l] = level[l];
    src.width[l] = width[l];
    src.pitch[l] = pitch[l];
  }
  return launch_corr1d_skew(src, num_levels, B, H, W1, skewed, skew_pitch, reinterpret_cast<cudaStream_t>(stream));
}

nnd_status nnd_corr1d_lookup_conv1x1_skewed(const float* const* skewed, const int* width, int skew_pitch, const float* coords,
                                            int B, int H, int W1, int num_levels, int radius, const float* weight,
                                            const float* bias, int c_out, int relu, int out_layout, void* out,
                                            nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(skewed && width && coords && weight && out, "lookup_conv1x1_skewed: null pointer argument");
  NND_REQUIRE(B > 0 && H > 0 && W1 > 0, "lookup_conv1x1_skewed: B, H, W1 must be positive");
  NND_REQUIRE(radius == 4 && num_levels == 4 && c_out == 256, "lookup_conv1x1_skewed: built for 4 levels, radius 4, 256 outputs");
  NND_REQUIRE(out_layout == 1 || out_layout == 2, "lookup_conv1x1_skewed: out_layout must be 1 (fp32) or 2 (fp16), channels-last");
  NND_REQUIRE(skew_pitch >= W1, "lookup_conv1x1_skewed: skew_pitch %d smaller than W1 %d", skew_pitch, W1);
  NND_REQUIRE(static_cast<long long>(H) * W1 < (1LL << 30), "lookup_conv1x1_skewed: H*W1 too large");
  LookupArgs a;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < num_levels; ++l) {
    NND_REQUIRE(width[l] >= 2 && skewed[l], "lookup_conv1x1_skewed: level %d invalid (linear_sampler needs width >= 2)", l);
    a.src[0].ptr[l] = skewed[l];
    a.src[0].width[l] = width[l];
    a.src[0].pitch[l] = skew_pitch;
  }
  a.coords = coords;
  a.out = reinterpret_cast<float*>(out);
  a.hw = H * W1;
  a.G = 1;
  a.n_src = 1;
  a.num_levels = num_levels;
  a.radius = radius;
  a.vec = 1;
  return launch_lookup_conv1x1_ws(a, weight, bias, relu ? 1 : 0, out_layout == 2 ? 1 : 0, static_cast<long long>(B) * a.hw, W1,
                                  reinterpret_cast<cudaStream_t>(stream));
}

nnd_status nnd_corr1d_lookup_skewed(const float* const* skewed, const int* width, int skew_pitch, const float* coords, int B,
                                    int H, int W1, int num_levels, int radius, float* out, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(skewed && width && coords && out, "lookup_skewed: null pointer argument");
  NND_REQUIRE(B > 0 && H > 0 && W1 > 0, "lookup_skewed: B, H, W1 must be positive");
  NND_REQUIRE(radius == 4 && num_levels >= 1 && num_levels <= NND_MAX_LEVELS, "lookup_skewed: built for radius 4 and 1..%d levels",
              NND_MAX_LEVELS);
  NND_REQUIRE(skew_pitch >= W1, "lookup_skewed: skew_pitch %d smaller than W1 %d", skew_pitch, W1);
  NND_REQUIRE(static_cast<long long>(H) * W1 < (1LL << 30) && B <= 65535, "lookup_skewed: H*W1 or B too large");
  LookupArgs a;
  memset(&a, 0, sizeof(a));
  for (int l = 0; l < num_levels; ++l) {
    NND_REQUIRE(width[l] >= 2 && skewed[l], "lookup_skewed: level %d invalid (linear_sampler needs width >= 2)", l);
    a.src[0].ptr[l] = skewed[l];
    a.src[0].width[l] = width[l];
    a.src[0].pitch[l] = skew_pitch;
  }
  a.coords = coords;
  a.out = out;
  a.hw = H * W1;
  a.G = 1;
  a.n_src = 1;
  a.num_levels = num_levels;
  a.radius = radius;
  return launch_lookup_skewed(a, B, H, W1, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
