// Tensor-core build of the RAFT-Stereo correlation pyramid: TMA -> shared memory (RN-rounded to
// TF32 in place) -> tcgen05.mma (kind::tf32, fp32 accumulators in TMEM) -> pooled 4-level epilogue ->
// TMA stores.  sm_100a only.
//
// Replaces CorrBlock1D.corr + CorrBlock1D.__init__ (nndepth/models/raft_stereo/cost_volume.py:55-61,
// :12-34): torch.matmul(f1^T, f2) / C**0.5 followed by four avg_pool1d passes.
//
// Per epipolar row (b,h) the product is D[m][n] = sum_c f1[b,c,h,m] * f2[b,c,h,n].  In NCHW both
// operands have the *spatial* index contiguous and the contraction index strided by H*W, i.e. both
// are "MN-major" in UMMA terms -- which tcgen05 supports for TF32, so the features are consumed where
// they lie (no transposes, no staging pass):
//
//   * shared-memory operand layout = the canonical MN-major "SW128 / 32-byte base" atom: boxes of
//     [KB rows of c][128 B of w] with the 32-byte chunks of a row XORed with row % 4 -- the ONLY layout
//     tcgen05 accepts for MN-major 32-bit operands (the plain 128B swizzle silently yields zeros).
//     Atoms stack along c (stride byte offset 512); a 128-row M tile is four boxes (leading byte
//     offset = box size), an N<=256 tile up to eight.
//   * one producer thread feeds a ring of stages with TMA boxes {32 w, 1 h, 32 c, 1 b}
//     (SWIZZLE_128B_ATOM_32B lands exactly that layout).  Measured on B200: these 4-KB boxes of 32
//     strided, 16-byte-aligned 128-byte rows stream at ~4 TB/s chip-wide; a cp.async (LSU) loader with
//     the same footprint was slower (its issue/handshake cost sits in the loader warps' critical path).
//   * eight rounder warps take each landed stage and round it to TF32 with round-to-nearest IN PLACE
//     before the tensor core sees it: tcgen05 kind::tf32 TRUNCATES the low 13 mantissa bits, a
//     systematic shrink of every product that costs 0.0126 px of final EPE at KITTI/32 iterations
//     against 0.0002 px with RN (tools/exp_epe.py).  full[s] (TMA) -> ready[s] (rounded) -> MMA.
//   * one elected thread issues tcgen05.mma M=128, N=roundup16(W2), K=8 per 8 channels; both M tiles
//     of a row share every B stage.  TMEM is a ring of 512 / n_cols tile slots, one M tile of fp32
//     accumulators each: with 160-column tiles (KITTI) that is three slots, so the next row's first tile
//     starts accumulating while the epilogue still drains the previous row, and its second tile as soon
//     as the previous row's first tile has been read; narrower volumes double-buffer whole rows.
//   * four epilogue warps, each on its own (no cross-warp barrier), read their TMEM lane quarter 32
//     columns at a time (tcgen05.ld 32x32b.x32), scale by 1/sqrt(C), pool in registers (a thread owns
//     one volume row -> 2/4/8-wide poolings are intra-thread, summed pairwise and halved like
//     avg_pool1d), park the four level tiles in swizzled shared memory and TMA-store them (lanes 0-3
//     issue one level each); the tensor maps clip ragged widths (156/78/39/19) and ragged row tiles.
//
// The volume is written once and never re-read.  At BASELINE shapes the kernel is HBM-bound
// (K = 256: ~25 flop/B, ridge > 200), so the pipeline is sized for bytes in flight, not MMA issue.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace nnd {

namespace {

constexpr int KB = 32;            // channels per pipeline stage (4 UMMA k-steps of 8)
constexpr int BOX_W = 32;         // floats per box row = 128 bytes = one swizzle span
constexpr int BOX_BYTES = KB * BOX_W * 4;
constexpr int TILE_M = 128;
constexpr int MAX_M = 2 * TILE_M; // volume rows per pass: two M tiles share every B stage
constexpr int MAX_N = 256;
constexpr int EPI_WARPS = 4;      // warps 0-3: epilogue (TMEM lane quarter = warp index)
constexpr int MMA_WARP = 4;       // warp 4: MMA issuer + TMEM owner
constexpr int LOADER_WARP0 = 5;   // warps 5-12: round the landed operand tiles to TF32 (RN) in place
constexpr int LOADER_WARPS = 8;
constexpr int TMA_WARP = LOADER_WARP0 + LOADER_WARPS;  // warp 13: TMA producer
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int NUM_THREADS = 32 * (TMA_WARP + 1);
constexpr int TMEM_COLS = 512;
constexpr int CHUNK = 32;         // volume columns per epilogue step
constexpr int MAX_STAGES = 6;
constexpr int MAX_TSLOTS = 16;     // 512 TMEM columns / 32
constexpr int BAR_BYTES = 512;     // 3 * MAX_STAGES + 2 * MAX_TSLOTS mbarriers + the TMEM base word

// epilogue staging, per warp and buffer set: level 0..3 tiles of 32 rows x {32,16,8,4} floats (7.5 KB, padded
// to 8 KB so that every set keeps the 1024-byte alignment the 128B swizzle pattern is anchored to)
constexpr int EPI_L0 = 32 * 32 * 4, EPI_L1 = 32 * 16 * 4, EPI_L2 = 32 * 8 * 4;  // level 3: 32 * 4 * 4 more
constexpr int EPI_WARP_SET = 8192;
constexpr int EPI_SET = EPI_WARPS * EPI_WARP_SET;  // one buffer set of all four epilogue warps

struct BuildParams {
  int C, W1, W2;
  int rows;            // B * H
  int H;
  int num_levels;      // 1..4 fused
  int n_chunks;        // ceil(W2 / 256): column chunks, one job each
  int m_passes;        // ceil(W1 / 256): passes of <= 2 M tiles per job
  int a_region_boxes;  // ceil(min(W1, 256) / 32)
  int b_region_boxes;  // ceil(min(W2, 256) / 32)
  int stages;
  int stage_bytes;     // (a_region_boxes + b_region_boxes) * BOX_BYTES
  int lt_mode;         // 1: "leftover transposed" layout (128 < W1 <= 160): see the kernel comment
  int lt_cols;         // TMEM columns of one job in lt_mode: n_cols + 64
  int n_tslots;        // tile slots in TMEM: 512 / slot_cols (one 128-row M tile of accumulators each)
  int slot_cols;       // b_region_boxes * 32
  float scale_div;
  float scale_inv;
  int scale_is_pow2;
  long long jobs;      // rows * n_chunks
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  int spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins == 1024) {  // watchdog: a protocol bug must fault, not hang the GPU
      spins = 0;
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) __trap();
    }
  } while (!done);
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint32_t rna_tf32(uint32_t x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(__uint_as_float(x)));
  return y;
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: rows of 128 bytes (64 fp16 of K), 8-row atoms of 1024 bytes
// stacked along M / N (stride byte offset 1024; the leading byte offset is not used by swizzled K-major layouts).
// A k-step of 16 fp16 advances the start address by 32 bytes inside the swizzle span.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
  return d;
}
// D = F32, A = B = F16, both K-major
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(TILE_M >> 4) << 24);
}

// UMMA shared-memory descriptor, MN-major, SWIZZLE_128B_BASE32B (bit layout of cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (stride between 32-element MN atoms)
//   [32,46) stride byte offset >> 4 (stride between 4-row K atoms) | [46,48) version = 1 | [61,64) layout = 1
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_32b(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;   // SWIZZLE_128B_BASE32B
  return d;
}

// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both MN-major.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(TILE_M >> 4) << 24);
}

// Byte offset, inside an operand region, of 16-byte chunk j (= columns 4j..4j+3 of the region's first
// column) of channel row r of a stage: box j/8, row r, 32-byte chunks XORed with r % 4.
__device__ __forceinline__ uint32_t chunk_offset(int j, int r) {
  const int jj = j & 7;
  return static_cast<uint32_t>((j >> 3) * BOX_BYTES + r * 128 + ((((jj >> 1) ^ (r & 3)) << 5) | ((jj & 1) << 4)));
}

// A job = one epipolar row x one chunk of <= 256 volume columns; its M tiles are processed in passes of
// <= 2 tiles, each pass accumulating over all channels into one TMEM slot.
struct JobGeom {
  int row, b, h;
  int n0, n_ext, n_cols, n_mma, b_boxes;
};

__device__ __forceinline__ JobGeom job_geom(const BuildParams& p, long long job) {
  JobGeom g;
  g.row = static_cast<int>(job / p.n_chunks);
  const int nc = static_cast<int>(job - static_cast<long long>(g.row) * p.n_chunks);
  g.b = g.row / p.H;
  g.h = g.row - g.b * p.H;
  g.n0 = nc * MAX_N;
  g.n_ext = min(p.W2 - g.n0, MAX_N);
  g.b_boxes = (g.n_ext + BOX_W - 1) / BOX_W;
  g.n_cols = g.b_boxes * BOX_W;  // TMEM columns per M tile
  g.n_mma = (g.n_ext + 15) & ~15;
  return g;
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
// IN_F16: the feature maps are fp16 channels-last (N, H, W, C) -- what a cuDNN fp16 encoder hands over.  A pixel's channels
// are contiguous, so both operands are K-major: TMA boxes {64 c, 32 w} land as 32 rows of 128 bytes in the plain
// SWIZZLE_128B layout, an epipolar row is ONE contiguous run of W * C * 2 bytes in memory (the NCHW form reads 256
// separate 624-byte segments per map and row), kind::f16 multiplies the fp16 values exactly (they are what TF32 would
// keep of them) and nothing needs rounding: the rounder warps retire at once and the MMAs wait on the TMA barrier.
template <bool IN_F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
corr1d_build_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ CUtensorMap map_l0, const __grid_constant__ CUtensorMap map_l1,
                         const __grid_constant__ CUtensorMap map_l2, const __grid_constant__ CUtensorMap map_l3,
                         const BuildParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x stage_bytes] [2 x EPI_SET] [barriers] [tmem base]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* epi = smem + static_cast<size_t>(p.stages) * p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + 2 * EPI_SET);
  // bars: full[s] (TMA landed), ready[s] (rounded to TF32), empty[s] (MMAs retired), tmem_full[slot], tmem_empty[slot]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 2 * MAX_TSLOTS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto ready_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto tmem_full_bar = [&](uint32_t slot) { return bar_base + 8u * (3 * MAX_STAGES + slot); };
  auto tmem_empty_bar = [&](uint32_t slot) { return bar_base + 8u * (3 * MAX_STAGES + MAX_TSLOTS + slot); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(ready_bar(s), LOADER_WARPS);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t slot = 0; slot < static_cast<uint32_t>(p.n_tslots); ++slot) {
      mbar_init(tmem_full_bar(slot), 1);
      mbar_init(tmem_empty_bar(slot), EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    prefetch_tmap(&map_l0);
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr int KBE = IN_F16 ? 64 : KB;             // channels per pipeline stage: 128 bytes of K either way
  const int k_blocks = (p.C + KBE - 1) / KBE;
  const uint32_t b_region = static_cast<uint32_t>(p.a_region_boxes) * BOX_BYTES;

  if (warp == TMA_WARP) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long job = blockIdx.x; job < p.jobs; job += gridDim.x) {
        const JobGeom g = job_geom(p, job);
        for (int pass = 0; pass < p.m_passes; ++pass) {
          const int m_start = pass * MAX_M;
          const int a_boxes = (min(p.W1 - m_start, MAX_M) + BOX_W - 1) / BOX_W;
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sbase = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes);
            mbar_expect_tx(full_bar(stage), static_cast<uint32_t>((a_boxes + g.b_boxes) * BOX_BYTES));
            for (int i = 0; i < a_boxes; ++i) {
              if (IN_F16) tma_load_4d(sbase + i * BOX_BYTES, &map_a, full_bar(stage), kb * KBE, m_start + i * BOX_W, g.h, g.b);
              else tma_load_4d(sbase + i * BOX_BYTES, &map_a, full_bar(stage), m_start + i * BOX_W, g.h, kb * KB, g.b);
            }
            for (int i = 0; i < g.b_boxes; ++i) {
              if (IN_F16) tma_load_4d(sbase + b_region + i * BOX_BYTES, &map_b, full_bar(stage), kb * KBE, g.n0 + i * BOX_W, g.h, g.b);
              else tma_load_4d(sbase + b_region + i * BOX_BYTES, &map_b, full_bar(stage), g.n0 + i * BOX_W, g.h, kb * KB, g.b);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= LOADER_WARP0 && IN_F16) {
    // fp16 operands are exact: nothing to round
  } else if (warp >= LOADER_WARP0) {
    // ===================== rounders: TF32 round-to-nearest of the landed stage, in place =====================
    // The tensor core truncates fp32 operands to TF32; rounding them to nearest first removes the bias.
    // Every thread owns a fixed set of 16-byte chunks of a stage (whatever the boxes hold), so no
    // geometry is needed here; stale chunks of partially filled regions are rounded harmlessly.
    const int wl = warp - LOADER_WARP0;
    constexpr int NCH = (KB / LOADER_WARPS) * 4;
    uint32_t off[NCH];
#pragma unroll
    for (int rr = 0; rr < KB / LOADER_WARPS; ++rr) {
      const int r = wl + LOADER_WARPS * rr;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int j = lane + 32 * s;
        off[rr * 4 + s * 2 + 0] = j < p.a_region_boxes * 8 ? chunk_offset(j, r) : 0xffffffffu;
        off[rr * 4 + s * 2 + 1] = j < p.b_region_boxes * 8 ? b_region + chunk_offset(j, r) : 0xffffffffu;
      }
    }
    int stage = 0;
    uint32_t phase = 0;
    long long iters = 0;
    for (long long job = blockIdx.x; job < p.jobs; job += gridDim.x) iters += static_cast<long long>(p.m_passes) * k_blocks;
    for (long long it = 0; it < iters; ++it) {
      mbar_wait(full_bar(stage), phase);
      uint8_t* sbase = smem + static_cast<size_t>(stage) * p.stage_bytes;
      {
        uint4 val[NCH];
        // all loads first, then the conversions, then the stores: independent chunks overlap their latencies
#pragma unroll
        for (int i = 0; i < NCH; ++i)
          val[i] = off[i] != 0xffffffffu ? *reinterpret_cast<const uint4*>(sbase + off[i]) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          val[i].x = rna_tf32(val[i].x); val[i].y = rna_tf32(val[i].y);
          val[i].z = rna_tf32(val[i].z); val[i].w = rna_tf32(val[i].w);
        }
#pragma unroll
        for (int i = 0; i < NCH; ++i)
          if (off[i] != 0xffffffffu) *reinterpret_cast<uint4*>(sbase + off[i]) = val[i];
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(ready_bar(stage));  // one arrival per rounder warp
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // operand descriptor of k-step ks of the region at `base`; the stage is ready when the rounders (fp32) or the TMA
      // itself (fp16) say so
      auto opdesc = [&](uint32_t base, int ks) -> uint64_t {
        return IN_F16 ? umma_desc_k_sw128(base + ks * 32) : umma_desc_mn_sw128_32b(base + ks * 1024, BOX_BYTES, 512);
      };
      auto stage_ready = [&](int st) -> uint32_t { return IN_F16 ? full_bar(st) : ready_bar(st); };
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if (IN_F16) tc_mma_f16(d, a, b, idesc, acc); else tc_mma_tf32(d, a, b, idesc, acc);
      };
      auto make_idesc = [&](int n) -> uint32_t { return IN_F16 ? umma_idesc_f16(n) : umma_idesc_tf32(n); };
      if (p.lt_mode) {
        // Leftover-transposed layout (one job = one row, 128 < W1 <= 160, one pass):
        //   D0  [cols 0, n_cols)             rows m = 0..127        x all n        (A = f1 boxes 0-3, B = f2)
        //   DLa [n_cols, n_cols + 32)        rows n = 0..127        x m' = 128+..  (A = f2 boxes 0-3, B = f1 box 4)
        //   DLb [n_cols + 32, n_cols + 64)   rows n = 128..255      x m'           (A = f2 boxes 4-7, B = f1 box 4)
        // The ragged second M tile (28 valid rows at KITTI) would otherwise occupy n_cols more columns; this way a
        // job needs n_cols + 64 columns and TWO jobs fit in TMEM: the epilogue of one row fully overlaps the
        // MMAs of the next.  Both operands are MN-major atoms of the same shape, so f2 boxes serve as the A operand
        // and the f1 box as B without any other change.
        const uint32_t idesc32 = make_idesc(32);
        uint32_t job_seq = 0;
        for (long long job = blockIdx.x; job < p.jobs; job += gridDim.x, ++job_seq) {
          const JobGeom g = job_geom(p, job);
          const uint32_t idesc = make_idesc(g.n_mma);
          const uint32_t slot = job_seq & 1;
          const uint32_t d0 = tmem_base + slot * p.lt_cols, dla = d0 + g.n_cols, dlb = dla + 32;
          const bool has_b = g.n_ext > TILE_M;
          mbar_wait(tmem_empty_bar(slot), ((job_seq >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(stage_ready(stage), phase);
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes);
            const uint32_t bbase = sbase + b_region;
#pragma unroll
            for (int ks = 0; ks < KB / 8; ++ks) {
              const uint32_t acc = (kb | ks) != 0 ? 1u : 0u;
              const uint64_t f1_t0 = opdesc(sbase, ks);
              const uint64_t f1_b4 = opdesc(sbase + 4 * BOX_BYTES, ks);
              const uint64_t f2_t0 = opdesc(bbase, ks);
              const uint64_t f2_t1 = opdesc(bbase + 4 * BOX_BYTES, ks);
              mma(d0, f1_t0, f2_t0, idesc, acc);
              mma(dla, f2_t0, f1_b4, idesc32, acc);
              if (has_b) mma(dlb, f2_t1, f1_b4, idesc32, acc);
            }
            tc_commit(empty_bar(stage));
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          tc_commit(tmem_full_bar(slot));
        }
      }
      uint32_t tile_seq = 0;  // every M tile takes the next TMEM tile slot of the ring
      for (long long job = p.lt_mode ? p.jobs : blockIdx.x; job < p.jobs; job += gridDim.x) {
        const JobGeom g = job_geom(p, job);
        const uint32_t idesc = make_idesc(g.n_mma);
        for (int pass = 0; pass < p.m_passes; ++pass) {
          const int m_start = pass * MAX_M;
          const int tiles = (min(p.W1 - m_start, MAX_M) + TILE_M - 1) / TILE_M;
          uint32_t slot[2], d_base[2];
          for (int t = 0; t < tiles; ++t) {
            slot[t] = (tile_seq + t) % p.n_tslots;
            d_base[t] = tmem_base + slot[t] * p.slot_cols;
          }
          for (int kb = 0; kb < k_blocks; ++kb) {
            mbar_wait(stage_ready(stage), phase);
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes);
            const uint32_t bbase = sbase + b_region;
#pragma unroll
            for (int ks = 0; ks < KB / 8; ++ks) {
              const uint64_t bdesc = opdesc(bbase, ks);
              for (int t = 0; t < tiles; ++t) {
                if ((kb | ks) == 0) {
                  // first write into this tile slot: the epilogue must have drained its previous tenant
                  mbar_wait(tmem_empty_bar(slot[t]), (((tile_seq + t) / p.n_tslots) & 1) ^ 1);
                  tc_fence_after();
                }
                const uint64_t adesc = opdesc(sbase + t * 4 * BOX_BYTES, ks);
                mma(d_base[t], adesc, bdesc, idesc, (kb | ks) != 0 ? 1u : 0u);
              }
            }
            tc_commit(empty_bar(stage));  // frees this smem stage when the MMAs above retire
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          for (int t = 0; t < tiles; ++t) tc_commit(tmem_full_bar(slot[t]));  // accumulators of this pass complete
          tile_seq += tiles;
        }
      }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quarter = warp) =====================
    // Each warp drains its own 32 volume rows independently (no cross-warp barrier): TMEM -> registers ->
    // scale + pool -> its own swizzled staging tiles -> TMA stores issued by lanes 0..3 (one level each).
    // A warp whose 32 rows lie past W1 (the ragged second M tile) skips the tile altogether.
    const int quarter = warp;
    uint32_t tile_seq = 0;
    int chunk_parity = 0;
    uint8_t* my_sets = epi + quarter * EPI_WARP_SET;  // + chunk_parity * EPI_SET
    const CUtensorMap* my_map = lane == 0 ? &map_l0 : lane == 1 ? &map_l1 : lane == 2 ? &map_l2 : &map_l3;
    const uint32_t my_level_off = lane == 0 ? 0u : lane == 1 ? EPI_L0 : lane == 2 ? EPI_L0 + EPI_L1 : EPI_L0 + EPI_L1 + EPI_L2;
    const bool store_lane = lane < p.num_levels && lane < 4;
    // stage one 32-row x 32-column chunk (rows = my lanes) and TMA-store its four level tiles
    auto stage_and_store = [&](float (&v)[32], int col0, int mrow, int vol_row) {
      uint8_t* set = my_sets + chunk_parity * EPI_SET;
      if (store_lane) tma_store_wait_read<1>();  // my store that last read this buffer set has drained
      __syncwarp();
      {
        float4* dst = reinterpret_cast<float4*>(set + lane * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      float l1[16], l2[8], l3[4];
#pragma unroll
      for (int i = 0; i < 16; ++i) l1[i] = pool2(v[2 * i], v[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < 8; ++i) l2[i] = pool2(l1[2 * i], l1[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) l3[i] = pool2(l2[2 * i], l2[2 * i + 1]);
      if (p.num_levels > 1) {
        float4* dst = reinterpret_cast<float4*>(set + EPI_L0 + lane * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j ^ ((lane >> 1) & 3)] = make_float4(l1[4 * j], l1[4 * j + 1], l1[4 * j + 2], l1[4 * j + 3]);
      }
      if (p.num_levels > 2) {
        float4* dst = reinterpret_cast<float4*>(set + EPI_L0 + EPI_L1 + lane * 32);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          dst[j ^ ((lane >> 2) & 1)] = make_float4(l2[4 * j], l2[4 * j + 1], l2[4 * j + 2], l2[4 * j + 3]);
      }
      if (p.num_levels > 3)
        *reinterpret_cast<float4*>(set + EPI_L0 + EPI_L1 + EPI_L2 + lane * 16) = make_float4(l3[0], l3[1], l3[2], l3[3]);
      fence_proxy_async();
      __syncwarp();
      if (store_lane) {
        const int col = col0 >> lane;  // first column of this chunk at my level
        if (col < (p.W2 >> lane)) tma_store_3d(my_map, smem_u32(set) + my_level_off, col, mrow, vol_row);
        tma_store_commit();
      }
      chunk_parity ^= 1;
    };
    // the same for a TRANSPOSED chunk: my lanes are 32 consecutive volume COLUMNS n0 + lane, v[j] is volume row
    // mrow + j.  Level 0 goes to shared memory column-wise; the pooled levels pair neighbouring lanes.
    auto stage_and_store_transposed = [&](float (&v)[32], int n0, int mrow, int vol_row) {
      constexpr unsigned FULL = 0xffffffffu;
      uint8_t* set = my_sets + chunk_parity * EPI_SET;
      if (store_lane) tma_store_wait_read<1>();
      __syncwarp();
      float* t0 = reinterpret_cast<float*>(set);
      float* t1 = reinterpret_cast<float*>(set + EPI_L0);
      float* t2 = reinterpret_cast<float*>(set + EPI_L0 + EPI_L1);
      float* t3 = reinterpret_cast<float*>(set + EPI_L0 + EPI_L1 + EPI_L2);
      const int c1 = lane >> 1, c2 = lane >> 2, c3 = lane >> 3;
#pragma unroll
      for (int j = 0; j < 32; ++j) {   // j = tile row (volume row mrow + j), lane = tile column
        const float x0 = v[j];
        t0[j * 32 + ((((lane >> 2) ^ (j & 7)) << 2) | (lane & 3))] = x0;
        const float x1 = pool2(x0, __shfl_xor_sync(FULL, x0, 1));       // valid on even lanes: columns (lane, lane+1)
        const float x2 = pool2(x1, __shfl_xor_sync(FULL, x1, 2));       // valid on lanes % 4 == 0
        const float x3 = pool2(x2, __shfl_xor_sync(FULL, x2, 4));       // valid on lanes % 8 == 0
        if (p.num_levels > 1 && (lane & 1) == 0) t1[j * 16 + ((((c1 >> 2) ^ ((j >> 1) & 3)) << 2) | (c1 & 3))] = x1;
        if (p.num_levels > 2 && (lane & 3) == 0) t2[j * 8 + ((((c2 >> 2) ^ ((j >> 2) & 1)) << 2) | (c2 & 3))] = x2;
        if (p.num_levels > 3 && (lane & 7) == 0) t3[j * 4 + c3] = x3;
      }
      fence_proxy_async();
      __syncwarp();
      if (store_lane) {
        const int col = n0 >> lane;
        if (col < (p.W2 >> lane)) tma_store_3d(my_map, smem_u32(set) + my_level_off, col, mrow, vol_row);
        tma_store_commit();
      }
      chunk_parity ^= 1;
    };
    auto scale32 = [&](float (&v)[32]) {
      if (p.scale_is_pow2) {  // x / 2^k == x * 2^-k exactly
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __fmul_rn(v[i], p.scale_inv);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __fdiv_rn(v[i], p.scale_div);
      }
    };
    if (p.lt_mode) {
      uint32_t job_seq = 0;
      for (long long job = blockIdx.x; job < p.jobs; job += gridDim.x, ++job_seq) {
        const JobGeom g = job_geom(p, job);
        const int n_chunks32 = g.n_cols / CHUNK;
        const uint32_t slot = job_seq & 1;
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t d0 = tmem_base + slot * p.lt_cols + lane_sel, dla = d0 + g.n_cols, dlb = dla + 32;
        const bool has_b = 128 + quarter * 32 < g.n_ext;   // my lane quarter holds valid columns n >= 128
        mbar_wait(tmem_full_bar(slot), (job_seq >> 1) & 1);
        tc_fence_after();
        for (int ch = 0; ch < n_chunks32; ++ch) {             // rows m = 32*quarter + lane of the first 128
          float v[32];
          tc_ld32(d0 + ch * CHUNK, v);
          scale32(v);
          stage_and_store(v, g.n0 + ch * CHUNK, quarter * 32, g.row);
        }
        {
          float v[32];                                          // volume rows 128 + j, columns n = 32*quarter + lane
          tc_ld32(dla, v);
          if (!has_b) {
            tc_fence_before();
            mbar_arrive(tmem_empty_bar(slot));                 // my last TMEM read of this job
          }
          scale32(v);
          stage_and_store_transposed(v, g.n0 + quarter * 32, TILE_M, g.row);
        }
        if (has_b) {
          float v[32];                                          // columns n = 128 + 32*quarter + lane
          tc_ld32(dlb, v);
          tc_fence_before();
          mbar_arrive(tmem_empty_bar(slot));
          scale32(v);
          stage_and_store_transposed(v, g.n0 + TILE_M + quarter * 32, TILE_M, g.row);
        }
      }
    }
    for (long long job = p.lt_mode ? p.jobs : blockIdx.x; job < p.jobs; job += gridDim.x) {
      const JobGeom g = job_geom(p, job);
      const int n_chunks32 = g.n_cols / CHUNK;
      for (int pass = 0; pass < p.m_passes; ++pass) {
        const int m_start = pass * MAX_M;
        const int tiles = (min(p.W1 - m_start, MAX_M) + TILE_M - 1) / TILE_M;
        const int last_tile = tiles - 1;
        for (int t = 0; t <= last_tile; ++t, ++tile_seq) {
          const uint32_t slot = tile_seq % p.n_tslots;
          const uint32_t d_base = tmem_base + slot * p.slot_cols + (static_cast<uint32_t>(quarter * 32) << 16);
          mbar_wait(tmem_full_bar(slot), (tile_seq / p.n_tslots) & 1);
          tc_fence_after();
          const int mrow = m_start + t * TILE_M + quarter * 32;
          if (mrow >= p.W1) {  // no valid rows of this tile in my lane quarter: hand the slot back at once
            tc_fence_before();
            mbar_arrive(tmem_empty_bar(slot));
            continue;
          }
          for (int ch = 0; ch < n_chunks32; ++ch) {
            float v[32];
            tc_ld32(d_base + ch * CHUNK, v);
            if (ch == n_chunks32 - 1) {
              // my last TMEM read of this tile: the slot goes back to the MMA warp while I finish the stores
              tc_fence_before();
              mbar_arrive(tmem_empty_bar(slot));
            }
            if (p.scale_is_pow2) {  // x / 2^k == x * 2^-k exactly
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __fmul_rn(v[i], p.scale_inv);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __fdiv_rn(v[i], p.scale_div);
            }

            uint8_t* set = my_sets + chunk_parity * EPI_SET;
            if (store_lane) tma_store_wait_read<1>();  // my store that last read this buffer set has drained
            __syncwarp();
            // level 0: 8 x 16-byte chunks per row, 128B swizzle (chunk ^= row & 7)
            {
              float4* dst = reinterpret_cast<float4*>(set + lane * 128);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                dst[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            float l1[16], l2[8], l3[4];
#pragma unroll
            for (int i = 0; i < 16; ++i) l1[i] = pool2(v[2 * i], v[2 * i + 1]);
#pragma unroll
            for (int i = 0; i < 8; ++i) l2[i] = pool2(l1[2 * i], l1[2 * i + 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) l3[i] = pool2(l2[2 * i], l2[2 * i + 1]);
            if (p.num_levels > 1) {  // 64 B per row, 64B swizzle (chunk ^= (row >> 1) & 3)
              float4* dst = reinterpret_cast<float4*>(set + EPI_L0 + lane * 64);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j ^ ((lane >> 1) & 3)] = make_float4(l1[4 * j], l1[4 * j + 1], l1[4 * j + 2], l1[4 * j + 3]);
            }
            if (p.num_levels > 2) {  // 32 B per row, 32B swizzle (chunk ^= (row >> 2) & 1)
              float4* dst = reinterpret_cast<float4*>(set + EPI_L0 + EPI_L1 + lane * 32);
#pragma unroll
              for (int j = 0; j < 2; ++j)
                dst[j ^ ((lane >> 2) & 1)] = make_float4(l2[4 * j], l2[4 * j + 1], l2[4 * j + 2], l2[4 * j + 3]);
            }
            if (p.num_levels > 3) {  // 16 B per row, no swizzle
              *reinterpret_cast<float4*>(set + EPI_L0 + EPI_L1 + EPI_L2 + lane * 16) =
                  make_float4(l3[0], l3[1], l3[2], l3[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (store_lane) {
              const int col = (g.n0 + ch * CHUNK) >> lane;  // first column of this chunk at my level
              if (col < (p.W2 >> lane)) tma_store_3d(my_map, smem_u32(set) + my_level_off, col, mrow, g.row);
              tma_store_commit();
            }
            chunk_parity ^= 1;
          }
        }
      }
    }
    // (the stores must have read their staging tiles before the CTA retires; their writes complete asynchronously and
    // are ordered by the end of the kernel)
    if (store_lane) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
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

nnd_status make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box, CUtensorMapSwizzle swizzle, const char* what,
                    CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT32) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("corr1d_build(tf32): cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return NND_ERR_CUDA;
  }
  cuuint32_t elem_strides[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims,
                  strides_bytes, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("corr1d_build(tf32): cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, static_cast<int>(r));
    return NND_ERR_CUDA;
  }
  return NND_OK;
}

}  // namespace

// in_f16 = 0: fp32 NCHW feature maps (TF32 products); 1: fp16 channels-last (N, H, W, C) feature maps (fp16 products)
static nnd_status corr1d_build_tc(const void* fmap1, const void* fmap2, int in_f16, int B, int C, int H, int W1, int W2,
                                  int num_levels, float* const* level, const int* pitch, cudaStream_t stream) {
  // TMA constraints: every global stride a multiple of 16 bytes, bases 16-byte aligned
  if (W1 % 4 != 0 || W2 % 4 != 0 || !aligned16(fmap1) || !aligned16(fmap2)) {
    set_error("corr1d_build(tf32): TMA needs W1 and W2 to be multiples of 4 and 16-byte aligned "
              "feature maps (W1=%d, W2=%d); use NND_PREC_FP32 for this shape", W1, W2);
    return NND_ERR_UNSUPPORTED;
  }
  if (in_f16 && C % 8 != 0) {
    set_error("corr1d_build(fp16 channels-last): TMA needs C to be a multiple of 8 (C=%d)", C);
    return NND_ERR_UNSUPPORTED;
  }
  for (int l = 0; l < num_levels; ++l) {
    if (pitch[l] % 4 != 0 || !aligned16(level[l])) {
      set_error("corr1d_build(tf32): level %d needs a pitch multiple of 4 floats and a 16-byte aligned base", l);
      return NND_ERR_UNSUPPORTED;
    }
  }
  BuildParams p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.W1 = W1; p.W2 = W2; p.H = H;
  p.rows = B * H;
  p.num_levels = num_levels;
  p.n_chunks = (W2 + MAX_N - 1) / MAX_N;
  p.m_passes = (W1 + MAX_M - 1) / MAX_M;
  p.jobs = static_cast<long long>(p.rows) * p.n_chunks;
  p.scale_div = static_cast<float>(sqrt(static_cast<double>(C)));
  p.scale_inv = 1.0f / p.scale_div;
  {
    int e = 0;
    p.scale_is_pow2 = (frexpf(p.scale_div, &e) == 0.5f) ? 1 : 0;
  }
  p.a_region_boxes = (min(W1, MAX_M) + BOX_W - 1) / BOX_W;
  p.b_region_boxes = (min(W2, MAX_N) + BOX_W - 1) / BOX_W;
  // an M tile always reads four boxes: keep the boxes behind a short A region inside the allocation
  const int a_tiles = (min(W1, MAX_M) + TILE_M - 1) / TILE_M;
  p.stage_bytes = (p.a_region_boxes + p.b_region_boxes) * BOX_BYTES;
  const int tail_boxes = 4 * a_tiles - p.a_region_boxes - p.b_region_boxes;  // > 0 only for tiny W2
  const int budget = 227 * 1024 - 1024 /*alignment slack*/ - 2 * EPI_SET - BAR_BYTES /*barriers*/ -
                     (tail_boxes > 0 ? tail_boxes * BOX_BYTES : 0);
  p.stages = budget / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.stages < 2) {
    set_error("corr1d_build(tf32): a pipeline stage of %d bytes leaves no room for double buffering", p.stage_bytes);
    return NND_ERR_UNSUPPORTED;
  }
  p.slot_cols = p.b_region_boxes * BOX_W;         // TMEM columns of one M tile
  p.n_tslots = TMEM_COLS / p.slot_cols;            // >= 2 (slot_cols <= 256): a ring of tile slots
  // leftover-transposed layout: a ragged second M tile of <= 32 rows, one column chunk, and room for two jobs
  p.lt_cols = p.slot_cols + 64;
  p.lt_mode = (W1 > TILE_M && W1 <= TILE_M + BOX_W && p.n_chunks == 1 && 2 * p.lt_cols <= TMEM_COLS) ? 1 : 0;
  { const char* e = getenv("NND_NO_LT"); if (e && e[0] == '1') p.lt_mode = 0; }
  const size_t smem_bytes = 1024 + static_cast<size_t>(p.stages) * p.stage_bytes + 2 * EPI_SET + BAR_BYTES;

  alignas(64) CUtensorMap map_a, map_b, map_l[4];
  if (in_f16) {
    // (N, H, W, C) fp16: dims innermost first {C, W, H, B}; a box is 64 channels (128 bytes) x 32 pixels of one row
    const cuuint64_t dims1[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W1), static_cast<cuuint64_t>(H),
                                 static_cast<cuuint64_t>(B)};
    const cuuint64_t str1[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W1) * C * 2,
                                static_cast<cuuint64_t>(H) * W1 * C * 2};
    const cuuint64_t dims2[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W2), static_cast<cuuint64_t>(H),
                                 static_cast<cuuint64_t>(B)};
    const cuuint64_t str2[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W2) * C * 2,
                                static_cast<cuuint64_t>(H) * W2 * C * 2};
    const cuuint32_t box[4] = {64, BOX_W, 1, 1};
    nnd_status st = make_map(&map_a, fmap1, 4, dims1, str1, box, CU_TENSOR_MAP_SWIZZLE_128B, "fmap1 (fp16 NHWC)",
                             CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    if (st != NND_OK) return st;
    st = make_map(&map_b, fmap2, 4, dims2, str2, box, CU_TENSOR_MAP_SWIZZLE_128B, "fmap2 (fp16 NHWC)",
                  CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    if (st != NND_OK) return st;
  } else {
    const cuuint64_t dims1[4] = {static_cast<cuuint64_t>(W1), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(C),
                                 static_cast<cuuint64_t>(B)};
    const cuuint64_t str1[3] = {static_cast<cuuint64_t>(W1) * 4, static_cast<cuuint64_t>(H) * W1 * 4,
                                static_cast<cuuint64_t>(C) * H * W1 * 4};
    const cuuint64_t dims2[4] = {static_cast<cuuint64_t>(W2), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(C),
                                 static_cast<cuuint64_t>(B)};
    const cuuint64_t str2[3] = {static_cast<cuuint64_t>(W2) * 4, static_cast<cuuint64_t>(H) * W2 * 4,
                                static_cast<cuuint64_t>(C) * H * W2 * 4};
    const cuuint32_t box[4] = {BOX_W, 1, KB, 1};
    nnd_status st = make_map(&map_a, fmap1, 4, dims1, str1, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "fmap1");
    if (st != NND_OK) return st;
    st = make_map(&map_b, fmap2, 4, dims2, str2, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "fmap2");
    if (st != NND_OK) return st;
  }
  const CUtensorMapSwizzle sw[4] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_32B,
                                    CU_TENSOR_MAP_SWIZZLE_NONE};
  for (int l = 0; l < 4; ++l) {
    const int ll = l < num_levels ? l : 0;  // unused maps alias level 0 (never dereferenced)
    const int w = W2 >> ll;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(W1),
                                static_cast<cuuint64_t>(p.rows)};
    const cuuint64_t str[2] = {static_cast<cuuint64_t>(pitch[ll]) * 4, static_cast<cuuint64_t>(W1) * pitch[ll] * 4};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(CHUNK >> l), 32, 1};
    nnd_status st = make_map(&map_l[l], level[ll], 3, dims, str, box, sw[l], "pyramid level");
    if (st != NND_OK) return st;
  }

  const long long sms = sm_count();  // persistent: one CTA per SM, rows dealt round-robin
  const unsigned grid = static_cast<unsigned>(p.jobs < sms ? p.jobs : sms);
  if (in_f16) {
    cudaError_t e = cudaFuncSetAttribute(corr1d_build_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(corr1d_build_tf32_kernel<fp16>)");
    corr1d_build_tf32_kernel<true><<<grid, NUM_THREADS, smem_bytes, stream>>>(map_a, map_b, map_l[0], map_l[1], map_l[2],
                                                                              map_l[3], p);
  } else {
    cudaError_t e = cudaFuncSetAttribute(corr1d_build_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(corr1d_build_tf32_kernel)");
    corr1d_build_tf32_kernel<false><<<grid, NUM_THREADS, smem_bytes, stream>>>(map_a, map_b, map_l[0], map_l[1], map_l[2],
                                                                               map_l[3], p);
  }
  return check_launch("corr1d_build_tf32_kernel");
}

nnd_status corr1d_build_tf32(const float* fmap1, const float* fmap2, int B, int C, int H, int W1, int W2,
                             int num_levels, float* const* level, const int* pitch, cudaStream_t stream) {
  return corr1d_build_tc(fmap1, fmap2, 0, B, C, H, W1, W2, num_levels, level, pitch, stream);
}

nnd_status corr1d_build_f16_nhwc(const void* fmap1, const void* fmap2, int B, int C, int H, int W1, int W2,
                                 int num_levels, float* const* level, const int* pitch, cudaStream_t stream) {
  return corr1d_build_tc(fmap1, fmap2, 1, B, C, H, W1, W2, num_levels, level, pitch, stream);
}

}  // namespace nnd
