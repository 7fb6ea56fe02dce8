// Warp-specialised tcgen05 form of the fused per-iteration kernel of RAFT-Stereo (sm_100a):
//
//     out[px, :] = relu(W . lookup(coords[px]) + b)          W = convc1's 1x1 weights, 36 -> 256
//
//   CorrBlock1D.__call__     raft_stereo/cost_volume.py:36-53   (4 levels x 9 taps, linear_sampler utils.py:4-27)
//   BasicMotionEncoder       blocks/update_block.py:51,58       cor = F.relu(self.convc1(corr))
//
// The (B, 36, H, W) lookup tensor never exists in HBM, and no warp ever waits for another role:
//
//   warps  0-15  PRODUCERS   warp = (32-pixel group of the 128-pixel tile, pyramid level); the four level-warps of a group
//                            sit on the four SM sub-partitions.  cp.async gathers the 32 windows of the NEXT tile into a
//                            private double buffer (4 lanes x 16 bytes per window, from a conservative 4-aligned start
//                            that needs no exact tap; coordinates prefetched two tiles ahead), interpolates the 9 taps of
//                            the current tile with the reference's exact fp32 operation order (bit-exact indices; one
//                            clamp per pixel, floor through FADD.RZ) and writes them, rounded to nearest TF32, straight
//                            into the A operand's core-matrix layout: K is laid out as 4 levels x 12 columns (9 taps +
//                            3 spare; level 0's spare columns carry 1.0 against the bias rows of the weight operand),
//                            so a thread's taps are three 16-byte stores.
//   warp   24    MMA         one thread: 6 x tcgen05.mma kind::tf32 (M 128 pixels, N 256 channels, K 8) per tile into
//                            one of two 256-column TMEM accumulators; tcgen05.commit releases the A slot to the
//                            producers and hands the accumulator to the epilogue.
//   warps 16-23  EPILOGUE    warp = (TMEM lane quarter, half of the 256 channels).  Per tile: tcgen05.ld 32 columns ->
//                            ReLU + fp16 in one conversion (or fp32) -> a 32-row x 128-byte box in the 128B-swizzle
//                            layout -> one TMA store (cp.async.bulk.tensor.2d) per box into the channels-last output: no
//                            read-back of the staging tile, no per-thread global stores.
//
// mbarrier pipelines: a_full / a_empty (producers <-> MMA, ring of NA operand slots), tmem_full / tmem_empty (MMA <->
// epilogue, 2 accumulators), b_ready (weights staged).  Work is split in CONTIGUOUS pixel ranges of equal length
// (one per SM, flat over the batch), so every CTA gathers and stores the same number of pixels: at the KITTI shape
// 405 pixels = 3.16 tiles each instead of 3 tiles on most SMs and 4 on 24 of them.
#include <cuda.h>

#include "common.cuh"
#include "lookup_common.cuh"

namespace nnd {
namespace ws {

constexpr int TILE = 128;                   // pixels per tile = UMMA M
constexpr int NOUT = 256;                   // output channels = UMMA N = TMEM columns per accumulator
constexpr int KCOLS = 12;                   // operand columns per level: 9 taps + 3 zeros
constexpr int KSTEPS = 6;                   // K = 48 in k-steps of 8
constexpr int A_KSTEP_BYTES = TILE * 32;
constexpr int B_KSTEP_BYTES = NOUT * 32;
constexpr int A_SLOT_BYTES = KSTEPS * A_KSTEP_BYTES;   // 24 KB
constexpr int NA = 2;                       // A-operand ring
constexpr int PROD_WARPS = 16;
constexpr int EPI_WARP0 = 16;
constexpr int EPI_WARPS = 8;
constexpr int MMA_WARP = 24;
constexpr int THREADS = 32 * (MMA_WARP + 1);
constexpr int WSTRIDE = 20;                 // floats per staged window (16 + 4: 16-byte aligned rows)
constexpr int WIN_BUF_FLOATS = 32 * WSTRIDE;
constexpr int BOX_ROW_BYTES = 128;          // one TMA store box: 32 pixel rows x 128 bytes (SWIZZLE_128B atom width)
constexpr int STAGE_WARP_BYTES = 32 * BOX_ROW_BYTES;

constexpr int SMEM_B = 0;
constexpr int SMEM_A = SMEM_B + KSTEPS * B_KSTEP_BYTES;                       // 48 KB
constexpr int SMEM_WIN = SMEM_A + NA * A_SLOT_BYTES;                          // + 48 KB
constexpr int SMEM_STAGE = SMEM_WIN + PROD_WARPS * 2 * WIN_BUF_FLOATS * 4;    // + 80 KB
constexpr int SMEM_BAR = SMEM_STAGE + EPI_WARPS * STAGE_WARP_BYTES;           // + 32 KB (1024-byte aligned boxes)
constexpr int SMEM_TOTAL = SMEM_BAR + 128;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Round-to-nearest (ties away) TF32 operand in ONE integer instruction: tcgen05.mma kind::tf32 ignores the low 13
// mantissa bits, so adding half an ulp of the 10-bit mantissa is all the rounding there is to do (cvt.rna.tf32.f32
// expands to a five-instruction sequence here).  Same value the MMA would see after cvt.rna for every finite input.
__device__ __forceinline__ uint32_t tf32_rna_bits(float x) { return __float_as_uint(x) + 0x1000u; }

__device__ __forceinline__ uint32_t pack_h2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// byte offset of the 16-byte chunk `ch` (4 consecutive K columns) of operand row `r`
__device__ __forceinline__ uint32_t chunk_offset(int r, int ch, int kstep_bytes) {
  return static_cast<uint32_t>((ch >> 1) * kstep_bytes + (r >> 3) * 256 + (ch & 1) * 128 + (r & 7) * 16);
}

}  // namespace ws

// out: channels-last (B*H*W, 256), fp16 (OUT_F16) or fp32.  weight: (36, 256) k-major fp32.
// SKEW: the pyramid is read from its SKEWED copy (corr1d_skew_kernel below): level l of an epipolar row (b, h) is stored
// as S[j][w1] with j = ((w1 >> l) - w2) mod W2_l, so the windows of neighbouring pixels with similar disparity lie in the
// SAME rows at neighbouring columns -- a warp's 32 windows are ~12 rows of 128 contiguous bytes instead of 32 separate
// 40-byte pieces in 32 volume rows (a 64-byte DRAM granule or two each).  a.src[0].pitch[l] is then the w1 pitch.
template <bool OUT_F16, bool SKEW>
__global__ void __launch_bounds__(ws::THREADS, 1)
corr1d_lookup_conv1x1_ws_kernel(const __grid_constant__ LookupArgs a, const __grid_constant__ CUtensorMap out_map,
                                const float* __restrict__ weight, const float* __restrict__ bias, int relu,
                                long long total_px, long long px_per_cta, int row_w1, unsigned magic_shl2) {
  using namespace ws;
  using umma::smem_u32;
  using umma::mbar_init;
  using umma::mbar_wait;
  constexpr int TAPS = 9, R = 4;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t bar0 = smem_u32(smem + SMEM_BAR);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (NA + s); };
  auto tmem_full = [&](int s) { return bar0 + 8u * (2 * NA + s); };
  auto tmem_empty = [&](int s) { return bar0 + 8u * (2 * NA + 2 + s); };
  const uint32_t b_ready = bar0 + 8u * (2 * NA + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 8 * (2 * NA + 5));

  const long long px0 = static_cast<long long>(blockIdx.x) * px_per_cta;
  const long long px_end = px0 + px_per_cta < total_px ? px0 + px_per_cta : total_px;
  const int nt = px_end > px0 ? static_cast<int>((px_end - px0 + TILE - 1) / TILE) : 0;

  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < NA; ++s) {
        mbar_init(a_full(s), PROD_WARPS);
        mbar_init(a_empty(s), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(tmem_full(s), 1);
        mbar_init(tmem_empty(s), 32 * EPI_WARPS);
      }
      mbar_init(b_ready, THREADS / 32);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // weights -> B operand (TF32, round to nearest), K laid out as 4 levels x (9 taps + 3 zero columns): staged by EVERY
  // thread (19 loads each, all issued before the first is used); the producers put their first gather in flight first.
  auto stage_weights = [&]() {
    constexpr int TOTAL = NOUT * KSTEPS * 8, PER_THREAD = (TOTAL + THREADS - 1) / THREADS;
    float wv[PER_THREAD];
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j) {
      const int idx = tid + THREADS * j, n = idx & (NOUT - 1), col = idx >> 8;
      const int l = col / KCOLS, k = col - l * KCOLS;
      // level 0's spare columns 9 and 10 carry the BIAS as two TF32 terms (the A operand holds 1.0 there): the tensor
      // core adds it, exact to 2**-22, and the epilogue needs neither the bias vector nor 256 additions per pixel
      const bool is_bias = l == 0 && (k == TAPS || k == TAPS + 1) && bias != nullptr;
      wv[j] = idx >= TOTAL ? 0.f : k < TAPS ? __ldg(weight + static_cast<long long>(l * TAPS + k) * NOUT + n)
                                  : is_bias ? __ldg(bias + n) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < PER_THREAD; ++j) {
      const int idx = tid + THREADS * j, n = idx & (NOUT - 1), col = idx >> 8;
      const int l = col / KCOLS, k = col - l * KCOLS;
      uint32_t bits = tf32_rna_bits(wv[j]);
      if (l == 0 && k == TAPS + 1)   // the remainder term of the bias
        bits = tf32_rna_bits(wv[j] - __uint_as_float((__float_as_uint(wv[j]) + 0x1000u) & 0xffffe000u));
      if (idx < TOTAL)
        *reinterpret_cast<uint32_t*>(smem + SMEM_B + chunk_offset(n, col >> 2, B_KSTEP_BYTES) + (col & 3) * 4) = bits;
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(b_ready);
  };

  if (warp < PROD_WARPS) {
    // =========================================== producers ===========================================
    // warp = 4 * (pixel group) + level: the four level-warps of one 32-pixel group sit on the four SM sub-partitions, so
    // a ragged last tile (one or two live groups) costs a quarter or half of a full one
    const int pg = warp >> 2, lvl = warp & 3;
    const int m = 32 * pg + lane;                       // my pixel's row of the tile
    const int w = a.src[0].width[lvl], pitch = a.src[0].pitch[lvl];
    const float* __restrict__ lbase = a.src[0].ptr[lvl];
    const LevelScale sc = level_scale(w, lvl, 0.f);
    const float inv_pow2 = 1.0f / static_cast<float>(1 << lvl);
    float* const wins = reinterpret_cast<float*>(smem + SMEM_WIN) + warp * (2 * WIN_BUF_FLOATS);
    const int q4 = 4 * (lane & 3), psub = lane >> 2;
    constexpr unsigned FULL = 0xffffffffu;

    auto coord_of = [&](int t) -> float {
      const long long px = px0 + static_cast<long long>(t) * TILE + m;
      return (t < nt && px < px_end) ? __ldg(a.coords + px) : 0.0f;
    };
    // Row layout: the window is fetched from a CONSERVATIVE 4-aligned start -- floor(centre - 4 - 1/64), clamped to the
    // row -- instead of the exact first tap index: never above it and at most one below (the taps' positions differ from
    // centre + dx by a few ulps only, launcher: width <= 16384), so with the alignment slack every tap lies in
    // [start, start + 14] of the 16 floats fetched, and the two exact taps the bounds used to cost are not computed.
    auto window_start = [&](float centre) -> int {
      return SKEW ? make_tap(0, R, centre, sc).i0
                  : (__float2int_rd(fminf(fmaxf(__fadd_rn(centre, -4.015625f), 0.f), sc.span)) & ~3);
    };
    constexpr int SKSTRIDE = 13;                         // skewed mode: 11 window floats per pixel, odd stride
    float sk[SKEW ? 11 : 1];                             // skewed mode: the next tile's window of this lane, in flight
    auto park_window = [&](float* win) {                 // skewed mode: registers -> my own window (no other lane reads it)
      if (SKEW) {
#pragma unroll
        for (int i = 0; i < 11; ++i) win[lane * SKSTRIDE + i] = sk[i];
        asm volatile("" ::: "memory");
      }
    };
    auto issue_windows = [&](int t, float c, int start, float* win) {
      const long long grp0 = px0 + static_cast<long long>(t) * TILE + 32 * pg;   // first pixel of my group
      if (t < nt && grp0 < px_end && SKEW) {
        const float centre = __fmul_rn(c, inv_pow2);
        const int lo = start;
        const int hi = make_tap(TAPS - 1, R, centre, sc).i1;
        const long long px = grp0 + lane;
        if (px < px_end) {
          const unsigned upx = static_cast<unsigned>(px);           // total_px < 2^31 (checked by the launcher)
          const long long bh = upx / static_cast<unsigned>(row_w1);
          const int w1 = static_cast<int>(upx - static_cast<unsigned>(bh) * static_cast<unsigned>(row_w1));
          const float* colp = lbase + bh * w * static_cast<long long>(pitch) + w1;   // row j = 0 of my epipolar row, my column
          int jr = ((w1 >> lvl) - lo) % w;               // row of window element 0; element i sits i rows above (mod w)
          if (jr < 0) jr += w;
          const float* src = colp + static_cast<long long>(jr) * pitch;
          const long long wrap = static_cast<long long>(w) * pitch;
          const int n_el = min(hi, w - 1) - lo;           // last window element
          // through REGISTERS (sk[], parked in the lane's window after this tile's taps): 4-byte cp.async copies of
          // scattered windows run at a third of the rate of plain loads
#pragma unroll
          for (int i = 0; i < 11; ++i) {
#ifndef NND_WS_SKIP_GATHER
            if (i <= n_el) sk[i] = __ldg(src);
#endif
            src = jr == 0 ? src + wrap - pitch : src - pitch;
            jr = jr == 0 ? w - 1 : jr - 1;
          }
        }
      } else if (t < nt && grp0 < px_end) {
        const float centre = __fmul_rn(c, inv_pow2);
        const int hi = __float2int_ru(fminf(fmaxf(__fadd_rn(centre, 4.015625f), 0.f), sc.span));   // >= the last tap's ceil
        const int n_live = static_cast<int>(px_end - grp0 < 32 ? px_end - grp0 : 32);
        const float* __restrict__ gbase = lbase + grp0 * pitch;
        const uint32_t dst0 = smem_u32(win + psub * WSTRIDE) + 4u * q4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = j * 8 + psub;
          const int cq = __shfl_sync(FULL, start, p) + q4;
          const int hp = __shfl_sync(FULL, hi, p);
#ifdef NND_WS_SKIP_GATHER
          if (false) {
#else
          if (p < n_live && cq <= hp) {
#endif
            const float* src = gbase + static_cast<unsigned>(p * pitch + cq);   // one IMAD.WIDE.U32
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + 4u * (j * 8 * WSTRIDE)), "l"(src) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float c0 = coord_of(0), c1 = coord_of(1);
    int s0 = window_start(__fmul_rn(c0, inv_pow2)), s1 = window_start(__fmul_rn(c1, inv_pow2));
    issue_windows(0, c0, s0, wins);
    stage_weights();
    park_window(wins);
    for (int t = 0; t < nt; ++t) {
      const float c2 = coord_of(t + 2);                          // in flight during this whole iteration
      issue_windows(t + 1, c1, s1, wins + ((t + 1) & 1) * WIN_BUF_FLOATS);
      asm volatile("cp.async.wait_group 1;" ::: "memory");         // tile t's windows (my copies) have landed
      __syncwarp();                                               // ... and the other lanes' copies
      const int slot = t % NA;
      mbar_wait(a_empty(slot), ((t / NA) & 1) ^ 1);               // the MMAs that read this slot have completed
      if (px0 + static_cast<long long>(t) * TILE + 32 * pg < px_end) {   // (rows past px_end are never stored)
        // centre clamped ONCE to [-6, span + 6] instead of every tap's x to [-1, span + 1]: inside the range nothing
        // changes, outside it every tap still lands on t = 0 or t = span through the saturation (NaN -> -6 -> t = 0)
        const float centre = fminf(fmaxf(__fmul_rn(c0, inv_pow2), -6.0f), __fadd_rn(sc.span, 6.0f));
        // byte address of window element 0, less the (2^23-as-float << 2) that the index trick below adds back
        // (magic_shl2 = 0x4b000000 << 2 arrives as a kernel argument: as a literal, ptxas splits it off the base again and
        // adds it back for every tap)
        const uint32_t mine = smem_u32(wins + (t & 1) * WIN_BUF_FLOATS + lane * (SKEW ? SKSTRIDE : WSTRIDE)) -
                              4u * static_cast<uint32_t>(s0) - magic_shl2;
        uint32_t v[12];
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
          // linear_sampler (utils.py:16-27) with the reference's operation order; t in [0, span]
          const float x = __fadd_rn(static_cast<float>(k - R), centre);
          const float qn = __fmul_rn(x, sc.inv_span);
          const float rr = __fmaf_rn(-qn, sc.span, x);
          const float tt = __fmul_rn(__saturatef(__fmaf_rn(rr, sc.inv_span, qn)), sc.span);
          const float u = __fadd_rz(tt, 8388608.0f);              // 2^23 + floor(t), exact: no F2I / FRND on the XU pipe
          const float f0 = __fadd_rn(u, -8388608.0f);
          const uint32_t addr = mine + (__float_as_uint(u) << 2);
          float v0, v1;
          asm volatile("ld.shared.f32 %0, [%2];\n\tld.shared.f32 %1, [%2+4];" : "=f"(v0), "=f"(v1) : "r"(addr) : "memory");
          const bool whole = (tt == f0);                          // ceil(t) == floor(t): both neighbours are element i0
          const float coef = whole ? 0.0f : __fsub_rn(__fadd_rn(f0, 1.0f), tt);   // coef = idx1 - t      (utils.py:26)
          const float one_minus = __fsub_rn(1.0f, coef);
          v1 = whole ? v0 : v1;
          // coef * val0 + (1 - coef) * val1, each operation rounded (utils.py:27); then RN to TF32 for the MMA
          v[k] = tf32_rna_bits(__fadd_rn(__fmul_rn(coef, v0), __fmul_rn(one_minus, v1)));
        }
        v[9] = v[10] = lvl == 0 ? 0x3f800000u : 0u;            // 1.0 x (bias_hi, bias_lo) rows of the B operand
        v[11] = 0u;
        unsigned char* slot_base = smem + SMEM_A + slot * A_SLOT_BYTES;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          *reinterpret_cast<uint4*>(slot_base + chunk_offset(m, 3 * lvl + i, A_KSTEP_BYTES)) =
              make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        fence_proxy_async();                                      // generic-proxy writes -> visible to the tensor core
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(slot));
      park_window(wins + ((t + 1) & 1) * WIN_BUF_FLOATS);
      c0 = c1;
      c1 = c2;
      s0 = s1;
      s1 = window_start(__fmul_rn(c2, inv_pow2));
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == MMA_WARP) {
    // =========================================== MMA issuer ===========================================
    stage_weights();
    if (lane == 0 && nt > 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(NOUT >> 3) << 17) |
                             (static_cast<uint32_t>(TILE >> 4) << 24);   // D f32, A/B tf32, both K-major
      mbar_wait(b_ready, 0);
      for (int t = 0; t < nt; ++t) {
        const int slot = t % NA, acc = t & 1;
        mbar_wait(tmem_empty(acc), ((t >> 1) & 1) ^ 1);           // the epilogue has drained this accumulator
        mbar_wait(a_full(slot), (t / NA) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          const uint64_t adesc = umma::desc_kmajor(smem_u32(smem + SMEM_A + slot * A_SLOT_BYTES + ks * A_KSTEP_BYTES), 128, 256);
          const uint64_t bdesc = umma::desc_kmajor(smem_u32(smem + SMEM_B + ks * B_KSTEP_BYTES), 128, 256);
          const uint32_t accum = ks > 0 ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem_base + static_cast<uint32_t>(acc * NOUT)), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
              : "memory");
        }
        tc_commit(a_empty(slot));        // implicit tcgen05.fence::before_thread_sync
        tc_commit(tmem_full(acc));
      }
    }
  } else {
    // =========================================== epilogue ===========================================
    const int e = warp - EPI_WARP0;
    const int eq = e & 3;                               // TMEM lane quarter this warp may read (warp % 4)
    const int half = e >> 2;                            // channels [128 * half, 128 * half + 128)
    if (e == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&out_map) : "memory");
    stage_weights();
    constexpr int ELEM = OUT_F16 ? 2 : 4;
    constexpr int ROW_BYTES = NOUT * ELEM;                       // bytes per pixel of the output
    constexpr int COLS_PER_BOX = BOX_ROW_BYTES / ELEM;           // 64 (fp16) or 32 (fp32) channels per TMA box
    constexpr int CHUNKS_PER_BOX = COLS_PER_BOX / 32;            // tcgen05.ld x32 chunks per box: 2 or 1
    constexpr int BOXES = (NOUT / 2) / COLS_PER_BOX;             // boxes per warp and tile: 2 or 4
    unsigned char* box = smem + SMEM_STAGE + e * STAGE_WARP_BYTES;          // [32 rows][128 B], 16-byte chunk c of row r at c ^ (r & 7)
    const uint32_t box_addr = smem_u32(box);
    unsigned char* out_bytes = reinterpret_cast<unsigned char*>(a.out);
    for (int t = 0; t < nt; ++t) {
      const int acc = t & 1;
      mbar_wait(tmem_full(acc), (t >> 1) & 1);
      tc_fence_after();
      const long long row0 = px0 + static_cast<long long>(t) * TILE + 32 * eq;   // pixel of my TMEM lane 0
      const bool whole = row0 + 32 <= px_end;            // all 32 rows are mine: TMA store; else (the global tail) plain stores
      if (row0 >= px_end) {                              // ragged last tile: none of my rows exists
        tc_fence_before();
        mbar_arrive(tmem_empty(acc));
        continue;
      }
#ifdef NND_WS_SKIP_EPI
      tc_fence_before();
      mbar_arrive(tmem_empty(acc));
      continue;
#endif
#pragma unroll 1
      for (int b = 0; b < BOXES; ++b) {
        const int colb = half * (NOUT / 2) + b * COLS_PER_BOX;   // first channel of the box
        if (whole) {
          if (lane == 0) tma_store_wait_read();          // the TMA store that last read this staging box has drained it
          __syncwarp();
        }
#pragma unroll
        for (int cq = 0; cq < CHUNKS_PER_BOX; ++cq) {
          const int col0 = colb + 32 * cq;
          float v[32];
          umma::ld32(tmem_base + (static_cast<uint32_t>(32 * eq) << 16) + static_cast<uint32_t>(acc * NOUT + col0), v);
          if (b == BOXES - 1 && cq == CHUNKS_PER_BOX - 1) {
            tc_fence_before();
            mbar_arrive(tmem_empty(acc));                        // my last TMEM read of this tile
          }
          // the bias is already in the accumulator (two extra K columns); fp16: ReLU rides on the conversion
          // (cvt.rn.relu: max(x, 0) then round == round then max, 0 is exact); fp32: one FMNMX per value
          uint4 pk[OUT_F16 ? 4 : 8];
          if (OUT_F16) {
            if (relu) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                pk[i] = make_uint4(pack_h2_relu(v[8 * i], v[8 * i + 1]), pack_h2_relu(v[8 * i + 2], v[8 * i + 3]),
                                   pack_h2_relu(v[8 * i + 4], v[8 * i + 5]), pack_h2_relu(v[8 * i + 6], v[8 * i + 7]));
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                pk[i] = make_uint4(pack_h2(v[8 * i], v[8 * i + 1]), pack_h2(v[8 * i + 2], v[8 * i + 3]),
                                   pack_h2(v[8 * i + 4], v[8 * i + 5]), pack_h2(v[8 * i + 6], v[8 * i + 7]));
            }
          } else {
            if (relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
              pk[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                                 __float_as_uint(v[4 * i + 3]));
          }
          constexpr int NPK = OUT_F16 ? 4 : 8;
          if (whole) {
#pragma unroll
            for (int i = 0; i < NPK; ++i)
              *reinterpret_cast<uint4*>(box + lane * BOX_ROW_BYTES + (((cq * NPK + i) ^ (lane & 7)) << 4)) = pk[i];
          } else if (row0 + lane < px_end) {
            uint4* dst = reinterpret_cast<uint4*>(out_bytes + (row0 + lane) * ROW_BYTES + static_cast<long long>(col0) * ELEM);
#pragma unroll
            for (int i = 0; i < NPK; ++i) dst[i] = pk[i];
          }
        }
        if (whole) {
          fence_proxy_async();                          // generic-proxy writes of the box -> visible to the TMA engine
          __syncwarp();
#ifndef NND_WS_SKIP_STORE
          if (lane == 0) {
            tma_store_2d(&out_map, box_addr, colb * ELEM / 4, static_cast<int>(row0));   // coordinates in 32-bit words, rows
            tma_store_commit();
          }
#endif
        }
      }
    }
    // the stores only have to have READ their staging boxes before the CTA retires; the writes themselves complete
    // asynchronously and are ordered by the end of the kernel (same contract as a CUTLASS TMA-store epilogue)
    if (lane == 0) tma_store_wait_read();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !sym) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// Skewed copy of a CorrBlock1D pyramid for the kernel above.  Source: level l rows (b, h, w1) of `pitch_l` floats
// (raft_stereo/cost_volume.py:29-34).  Destination: for every epipolar row (b, h) a (W2_l x P1) matrix
//     S_l[(b, h)][j][w1] = V_l[(b, h, w1)][w2],   j = ((w1 >> l) - w2) mod W2_l,   P1 = roundup4(W1).
// Block = 32 consecutive w1 of one (b, h) and level: their volume rows are staged in shared memory with coalesced
// loads, then every warp writes rows j as 128-byte segments.
// ------------------------------------------------------------------------------------------------
struct SkewArgs {
  ConstPyramid src;
  float* dst[NND_MAX_LEVELS];
  int W1, P1;
};

__global__ void __launch_bounds__(256)
corr1d_skew_kernel(const __grid_constant__ SkewArgs a) {
  extern __shared__ float rows[];                    // [32][WP]
  const int l = blockIdx.z;
  const int w = a.src.width[l], pitch = a.src.pitch[l];
  const int WP = l == 0 ? ((w + 1) & ~1) : (w | 1);  // the read below strides WP + 2^-l banks per lane: keep that odd
  const long long bh = blockIdx.y;
  const int w1_0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* __restrict__ src = a.src.ptr[l] + (bh * a.W1 + w1_0) * pitch;
  for (int r = warp; r < 32; r += 8) {
    if (w1_0 + r >= a.W1) break;
    for (int c = lane; c < w; c += 32) rows[r * WP + c] = __ldg(src + static_cast<long long>(r) * pitch + c);
  }
  __syncthreads();
  const int w1 = w1_0 + lane;
  if (w1 >= a.W1) return;
  float* __restrict__ dst = a.dst[l] + bh * w * static_cast<long long>(a.P1) + w1;
  int w2 = ((w1 >> l) - warp) % w;                   // row j = warp; every step of 8 rows moves w2 back by 8 (mod w)
  if (w2 < 0) w2 += w;
  const int step = 8 % w;
  const float* mine = rows + lane * WP;
  for (int j = warp; j < w; j += 8) {
    dst[static_cast<long long>(j) * a.P1] = mine[w2];
    w2 -= step;
    if (w2 < 0) w2 += w;
  }
}

nnd_status launch_corr1d_skew(const ConstPyramid& src, int num_levels, int B, int H, int W1, float* const* dst, int P1,
                              cudaStream_t stream) {
  SkewArgs a;
  memset(&a, 0, sizeof(a));
  a.src = src;
  a.W1 = W1;
  a.P1 = P1;
  int wmax = 0;
  for (int l = 0; l < num_levels; ++l) {
    a.dst[l] = dst[l];
    wmax = src.width[l] > wmax ? src.width[l] : wmax;
  }
  const size_t smem = static_cast<size_t>(32) * (wmax + 2) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("corr1d_skew: volume rows of %d floats do not fit shared memory", wmax);
    return NND_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024) cudaFuncSetAttribute(corr1d_skew_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const long long bh = static_cast<long long>(B) * H;
  if (bh > 65535) {
    set_error("corr1d_skew: B*H = %lld exceeds the grid limit (65535)", bh);
    return NND_ERR_INVALID_ARGUMENT;
  }
  dim3 grid((W1 + 31) / 32, static_cast<unsigned>(bh), num_levels);
  corr1d_skew_kernel<<<grid, 256, smem, stream>>>(a);
  return check_launch("corr1d_skew_kernel");
}

// ------------------------------------------------------------------------------------------------
// CorrBlock1D.__call__ (raft_stereo/cost_volume.py:36-53) on the SKEWED copy: warp = 32 consecutive pixels x one level.
// A lane fetches the <= 11 window elements of its pixel -- element e lies e rows above row j0 = ((w1 >> l) - i0) mod W2_l,
// in column w1 -- so for a smooth disparity field a warp load is one 128-byte row segment.  Same tap arithmetic as the
// row-layout kernel (bit-identical output), same (B, L*9, H, W) fp32 channel-plane stores.
// ------------------------------------------------------------------------------------------------
template <int TAPS>
__global__ void __launch_bounds__(32 * NND_MAX_LEVELS, 8)
corr1d_lookup_skewed_kernel(const __grid_constant__ LookupArgs a, int row_w1, int rows_h, unsigned magic_shl2) {
  constexpr int R = (TAPS - 1) / 2, NEL = TAPS + 3, STRIDE = NEL + 1;   // 12 window floats + 1 spare; odd stride: conflict-free
  extern __shared__ float lk_win[];
  const int lane = threadIdx.x, lvl = threadIdx.y;
  // block = 32 consecutive pixels of ONE epipolar row (b, h): blockIdx.y = b * H + h, so no per-thread division
  const int w1 = blockIdx.x * 32 + lane;
  if (w1 >= row_w1) return;                             // every lane works on its own window: no warp-wide step below
  const long long bh = blockIdx.y;
  const int b = blockIdx.y / rows_h, h = blockIdx.y - b * rows_h;
  const int rem = h * row_w1 + w1;
  float* win = lk_win + (lvl * 32 + lane) * STRIDE;
  const int w = a.src[0].width[lvl], pitch = a.src[0].pitch[lvl];
  const float c = __ldg(a.coords + bh * row_w1 + w1);
  const LevelScale sc = level_scale(w, lvl, 0.f);
  const float centre = __fmul_rn(c, 1.0f / static_cast<float>(1 << lvl));
  // conservative window [lo, lo + 11]: lo = floor(centre - 4 - 1/64) clamped to the row is never above the first tap's
  // index and at most one below it (see the fused kernel above; the launcher bounds the width)
  const int lo = __float2int_rd(fminf(fmaxf(__fadd_rn(centre, -4.015625f), 0.f), sc.span));
  const int hi = __float2int_ru(fminf(fmaxf(__fadd_rn(centre, 4.015625f), 0.f), sc.span));
  int jr = (w1 >> lvl) - lo;                            // in (-W2, W1 >> l]: one add, and a modulo only when W1 > W2
  if (jr < 0) jr += w;
  if (jr >= w) jr %= w;
  const float* colp = a.src[0].ptr[lvl] + bh * w * static_cast<long long>(pitch) + w1;   // row j = 0 of my epipolar row, my column
  unsigned off = static_cast<unsigned>(jr) * static_cast<unsigned>(pitch);             // 32-bit offsets inside the (W2 x P1) slab
  const unsigned top = static_cast<unsigned>(w - 1) * static_cast<unsigned>(pitch);
  const int n_el = hi - lo;
  // through registers: all twelve loads of a lane in flight at once (4-byte cp.async here was three times slower on
  // scattered windows, and a persistent, software-pipelined form of this kernel lost to plain oversubscription: 49.6 vs
  // 40.9 us at batch 64; ncu: issue slots 78 % busy -- the kernel is instruction-bound, 94 MB of DRAM reads)
  float v[NEL];
#pragma unroll
  for (int i = 0; i < NEL; ++i) {
    v[i] = i <= n_el ? __ldg(colp + off) : 0.f;
    off = off == 0 ? top : off - pitch;                  // element i + 1 lies one row above (mod W2)
  }
#pragma unroll
  for (int i = 0; i < NEL; ++i) win[i] = v[i];
  // taps: the reference's operation order (utils.py:16-27) with one clamp per pixel and floor through FADD.RZ, exactly
  // as in the fused kernel's producers
  const float cc = fminf(fmaxf(centre, -6.0f), __fadd_rn(sc.span, 6.0f));
  const uint32_t mine = static_cast<uint32_t>(__cvta_generic_to_shared(win)) - 4u * static_cast<uint32_t>(lo) - magic_shl2;
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
    const float coef = whole ? 0.0f : __fsub_rn(__fadd_rn(f0, 1.0f), tt);
    const float one_minus = __fsub_rn(1.0f, coef);
    v1 = whole ? v0 : v1;
    // coef * val0 + (1 - coef) * val1, each operation rounded (utils.py:27)
    *op = __fadd_rn(__fmul_rn(coef, v0), __fmul_rn(one_minus, v1));
    op += a.hw;
  }
}

nnd_status launch_lookup_skewed(const LookupArgs& a, int B, int H, int W1, cudaStream_t stream) {
  if (a.src[0].width[0] > 16384) {   // the conservative window start assumes positions exact to well below 1/64
    set_error("lookup_skewed: level-0 width %d exceeds 16384", a.src[0].width[0]);
    return NND_ERR_INVALID_ARGUMENT;
  }
  if (static_cast<long long>(B) * H > 65535) {
    set_error("lookup_skewed: B*H = %lld exceeds the grid limit (65535)", static_cast<long long>(B) * H);
    return NND_ERR_INVALID_ARGUMENT;
  }
  dim3 grid((W1 + 31) / 32, static_cast<unsigned>(B * H));
  dim3 block(32, a.num_levels);
  const size_t smem = static_cast<size_t>(a.num_levels) * 32 * 13 * sizeof(float);
  corr1d_lookup_skewed_kernel<9><<<grid, block, smem, stream>>>(a, W1, H, 0x4b000000u << 2);
  return check_launch("corr1d_lookup_skewed_kernel");
}

// host-side launcher, called by nnd_corr1d_lookup_conv1x1[_skewed] (lookup.cu) for the shipping shape:
// 4 levels, radius 4, c_out = 256, channels-last output, 16-byte aligned pyramid rows
nnd_status launch_lookup_conv1x1_ws(const LookupArgs& a, const float* weight, const float* bias, int relu, int out_f16,
                                    long long total_px, int skew_w1, cudaStream_t stream) {
  const long long sms = sm_count();
  const long long tiles = (total_px + ws::TILE - 1) / ws::TILE;
  const long long grid = tiles < sms ? tiles : sms;
  long long per = (total_px + grid - 1) / grid;
  per = (per + 31) & ~31LL;                                  // whole 32-pixel groups per CTA
  if (total_px >= (1LL << 31)) {
    set_error("lookup_conv1x1: %lld pixels exceed the TMA store's 32-bit row coordinate", total_px);
    return NND_ERR_INVALID_ARGUMENT;
  }
  if (a.src[0].width[0] > 16384) {   // the conservative window start assumes positions exact to well below 1/64
    set_error("lookup_conv1x1: level-0 width %d exceeds 16384", a.src[0].width[0]);
    return NND_ERR_INVALID_ARGUMENT;
  }
  // the channels-last output as a 2-D tensor of 32-bit words: (total_px rows) x (row_bytes / 4 words), stored in boxes of
  // 32 rows x 32 words (128 bytes) whose shared-memory image uses the 128-byte swizzle
  EncodeTiledFn enc = encode_tiled();
  if (!enc) {
    set_error("lookup_conv1x1: cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return NND_ERR_CUDA;
  }
  alignas(64) CUtensorMap out_map;
  const cuuint64_t row_bytes = static_cast<cuuint64_t>(ws::NOUT) * (out_f16 ? 2 : 4);
  const cuuint64_t dims[2] = {row_bytes / 4, static_cast<cuuint64_t>(total_px)};
  const cuuint64_t strides[1] = {row_bytes};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = enc(&out_map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, a.out, dims, strides, box, elem_strides,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("lookup_conv1x1: cuTensorMapEncodeTiled(output) failed with CUresult %d", static_cast<int>(r));
    return NND_ERR_CUDA;
  }
#define NND_WS_LAUNCH(F16, SK)                                                                                          \
  do {                                                                                                                  \
    cudaError_t e = cudaFuncSetAttribute(corr1d_lookup_conv1x1_ws_kernel<F16, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         ws::SMEM_TOTAL);                                                               \
    if (e != cudaSuccess) return cuda_fail(e, "lookup_conv1x1: shared-memory attribute");                               \
    corr1d_lookup_conv1x1_ws_kernel<F16, SK><<<static_cast<unsigned>(grid), ws::THREADS, ws::SMEM_TOTAL, stream>>>(     \
        a, out_map, weight, bias, relu, total_px, per, skew_w1 > 0 ? skew_w1 : 1, 0x4b000000u << 2);                                      \
  } while (0)
  if (skew_w1 > 0) {
    if (out_f16) NND_WS_LAUNCH(true, true); else NND_WS_LAUNCH(false, true);
  } else {
    if (out_f16) NND_WS_LAUNCH(true, false); else NND_WS_LAUNCH(false, false);
  }
#undef NND_WS_LAUNCH
  return check_launch("corr1d_lookup_conv1x1_ws_kernel");
}

}  // namespace nnd
