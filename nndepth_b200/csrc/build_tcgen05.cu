// Tensor-core build of the RAFT-Stereo correlation pyramid: TMA -> shared memory -> tcgen05.mma
// (kind::tf32, fp32 accumulators in TMEM) -> pooled 4-level epilogue -> TMA stores.  sm_100a only.
//
// Replaces CorrBlock1D.corr + CorrBlock1D.__init__ (nndepth/models/raft_stereo/cost_volume.py:55-61,
// :12-34): torch.matmul(f1^T, f2) / C**0.5 followed by four avg_pool1d passes.
//
// Per epipolar row (b,h) the product is D[m][n] = sum_c f1[b,c,h,m] * f2[b,c,h,n].  In NCHW both
// operands have the *spatial* index contiguous and the contraction index strided by H*W, i.e. both
// are "MN-major" in UMMA terms -- which tcgen05 supports for TF32, so the features are consumed where
// they lie (no transposes, no staging pass):
//
//   * TMA boxes {32 w, 1 h, KB c, 1 b} with SWIZZLE_128B_ATOM_32B land as [KB rows of c][128 B of w]:
//     exactly the canonical MN-major "SW128 / 32-byte base" atom (4 rows x 128 B, 32-byte chunks XORed
//     with row % 4) -- the ONLY shared-memory layout tcgen05 accepts for MN-major 32-bit operands (the
//     plain 128B swizzle silently yields zeros).  Atoms stack along c (stride byte offset 512); a
//     128-row M tile is four boxes (leading byte offset = box size), an N<=256 tile up to eight.
//   * one elected thread issues tcgen05.mma M=128, N=roundup16(W2), K=8 per 8 channels; the whole
//     W1 x W2 row (W1, W2 <= 256 per job) accumulates in TMEM: two M tiles x <=256 columns = 512.
//   * four epilogue warps read their TMEM lane quarter 32 columns at a time (tcgen05.ld 32x32b.x32),
//     scale by 1/sqrt(C), pool in registers (a thread owns one volume row -> 2/4/8-wide poolings are
//     intra-thread, summed pairwise and halved like avg_pool1d), park the four level tiles in
//     swizzled shared memory and TMA-store them; the tensor maps clip ragged widths (156/78/39/19).
//
// The volume is written once and never re-read.  At BASELINE shapes the kernel is HBM-bound
// (K = 256: ~25 flop/B, ridge > 200), so the pipeline is sized for bytes in flight, not MMA issue.
#include <cuda.h>
#include <math.h>

#include "common.cuh"

namespace nnd {

namespace {

constexpr int KB = 32;            // channels per pipeline stage (4 UMMA k-steps of 8)
constexpr int BOX_W = 32;         // floats per TMA box row = 128 bytes = one swizzle span
constexpr int BOX_BYTES = KB * BOX_W * 4;
constexpr int TILE_M = 128;
constexpr int MAX_N = 256;
constexpr int NUM_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int EPI_THREADS = 128;
constexpr int TMEM_COLS = 512;
constexpr int CHUNK = 32;         // volume columns per epilogue step

// epilogue staging (one buffer set): level 0..3 tiles of 128 rows x {32,16,8,4} floats
constexpr int EPI_L0 = 128 * 32 * 4, EPI_L1 = 128 * 16 * 4, EPI_L2 = 128 * 8 * 4, EPI_L3 = 128 * 4 * 4;
constexpr int EPI_SET = EPI_L0 + EPI_L1 + EPI_L2 + EPI_L3;  // 30720 B

struct BuildParams {
  int C, W1, W2;
  int rows;         // B * H
  int H;
  int num_levels;   // 1..4 fused
  int m_groups;     // ceil(W1 / 256)
  int n_chunks;     // ceil(W2 / 256)
  int stages;
  int stage_bytes;  // (a_boxes_padded + b_boxes) * BOX_BYTES
  int a_boxes_pad;  // roundup4(max boxes of an m-group)
  float scale_div;
  float scale_inv;
  int scale_is_pow2;
  long long jobs;   // rows * m_groups * n_chunks
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
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
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

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

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
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
  // bars[0..S) full, [S..2S) empty, [2S] tmem_full, [2S+1] tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_empty_bar = bar_base + 8u * (2 * p.stages + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    prefetch_tmap(&map_l0);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int k_blocks = (p.C + KB - 1) / KB;
  const int jobs_per_row = p.m_groups * p.n_chunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long job = blockIdx.x; job < p.jobs; job += gridDim.x) {
        const int row = static_cast<int>(job / jobs_per_row);
        const int sub = static_cast<int>(job - static_cast<long long>(row) * jobs_per_row);
        const int mg = sub / p.n_chunks, nc = sub - mg * p.n_chunks;
        const int b = row / p.H, h = row - b * p.H;
        const int m0 = mg * 2 * TILE_M, n0 = nc * MAX_N;
        const int a_boxes = (min(p.W1 - m0, 2 * TILE_M) + BOX_W - 1) / BOX_W;
        const int b_boxes = (min(p.W2 - n0, MAX_N) + BOX_W - 1) / BOX_W;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sbase = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes);
          mbar_expect_tx(full_bar(stage), static_cast<uint32_t>((a_boxes + b_boxes) * BOX_BYTES));
          for (int i = 0; i < a_boxes; ++i)
            tma_load_4d(sbase + i * BOX_BYTES, &map_a, full_bar(stage), m0 + i * BOX_W, h, kb * KB, b);
          const uint32_t bbase = sbase + p.a_boxes_pad * BOX_BYTES;
          for (int i = 0; i < b_boxes; ++i)
            tma_load_4d(bbase + i * BOX_BYTES, &map_b, full_bar(stage), n0 + i * BOX_W, h, kb * KB, b);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long job = blockIdx.x; job < p.jobs; job += gridDim.x) {
        const int row = static_cast<int>(job / jobs_per_row);
        const int sub = static_cast<int>(job - static_cast<long long>(row) * jobs_per_row);
        const int mg = sub / p.n_chunks, nc = sub - mg * p.n_chunks;
        const int m_ext = min(p.W1 - mg * 2 * TILE_M, 2 * TILE_M);
        const int n_ext = min(p.W2 - nc * MAX_N, MAX_N);
        const int m_tiles = (m_ext + TILE_M - 1) / TILE_M;
        const int n_mma = (n_ext + 15) & ~15;
        const int n_cols = ((n_ext + BOX_W - 1) / BOX_W) * BOX_W;  // TMEM columns per M tile
        const uint32_t idesc = umma_idesc_tf32(n_mma);
        mbar_wait(tmem_empty_bar, acc_phase ^ 1);  // epilogue has drained the previous job
        tc_fence_after();
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + static_cast<size_t>(stage) * p.stage_bytes);
          const uint32_t bbase = sbase + p.a_boxes_pad * BOX_BYTES;
#pragma unroll
          for (int ks = 0; ks < KB / 8; ++ks) {
            const uint64_t bdesc = umma_desc_mn_sw128_32b(bbase + ks * 1024, BOX_BYTES, 512);
            for (int mt = 0; mt < m_tiles; ++mt) {
              const uint64_t adesc = umma_desc_mn_sw128_32b(sbase + mt * 4 * BOX_BYTES + ks * 1024, BOX_BYTES, 512);
              tc_mma_tf32(tmem_base + mt * n_cols, adesc, bdesc, idesc, (kb | ks) != 0 ? 1u : 0u);
            }
          }
          tc_commit(empty_bar(stage));  // frees this smem stage when the MMAs above retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(tmem_full_bar);  // accumulators complete
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quarter = warp % 4) =====================
    const int quarter = warp & 3;
    const int trow = quarter * 32 + lane;  // TMEM lane == row of the M tile
    const int et = threadIdx.x - 64;       // 0..127 among the epilogue threads
    uint32_t acc_phase = 0;
    int chunk_parity = 0;
    for (long long job = blockIdx.x; job < p.jobs; job += gridDim.x) {
      const int row = static_cast<int>(job / jobs_per_row);
      const int sub = static_cast<int>(job - static_cast<long long>(row) * jobs_per_row);
      const int mg = sub / p.n_chunks, nc = sub - mg * p.n_chunks;
      const int m0 = mg * 2 * TILE_M, n0 = nc * MAX_N;
      const int m_ext = min(p.W1 - m0, 2 * TILE_M);
      const int n_ext = min(p.W2 - n0, MAX_N);
      const int m_tiles = (m_ext + TILE_M - 1) / TILE_M;
      const int n_cols = ((n_ext + BOX_W - 1) / BOX_W) * BOX_W;
      const int n_chunks32 = n_cols / CHUNK;

      mbar_wait(tmem_full_bar, acc_phase);
      tc_fence_after();
      for (int mt = 0; mt < m_tiles; ++mt) {
        for (int ch = 0; ch < n_chunks32; ++ch) {
          float v[32];
          tc_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + mt * n_cols + ch * CHUNK, v);
          if (mt == m_tiles - 1 && ch == n_chunks32 - 1) {
            // last TMEM read of this job: hand the accumulators back to the MMA warp
            tc_fence_before();
            mbar_arrive(tmem_empty_bar);
          }
          if (p.scale_is_pow2) {  // x / 2^k == x * 2^-k exactly
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __fmul_rn(v[i], p.scale_inv);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __fdiv_rn(v[i], p.scale_div);
          }

          uint8_t* set = epi + chunk_parity * EPI_SET;
          if (et == 0) tma_store_wait_read<1>();  // the store that last read this buffer set has drained
          epi_bar_sync();
          // level 0: 8 x 16-byte chunks per row, 128B swizzle (chunk ^= row & 7)
          {
            float4* dst = reinterpret_cast<float4*>(set + trow * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j ^ (trow & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          float l1[16], l2[8], l3[4];
#pragma unroll
          for (int i = 0; i < 16; ++i) l1[i] = pool2(v[2 * i], v[2 * i + 1]);
#pragma unroll
          for (int i = 0; i < 8; ++i) l2[i] = pool2(l1[2 * i], l1[2 * i + 1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) l3[i] = pool2(l2[2 * i], l2[2 * i + 1]);
          if (p.num_levels > 1) {  // 64 B per row, 64B swizzle (chunk ^= (row >> 1) & 3)
            float4* dst = reinterpret_cast<float4*>(set + EPI_L0 + trow * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j ^ ((trow >> 1) & 3)] = make_float4(l1[4 * j], l1[4 * j + 1], l1[4 * j + 2], l1[4 * j + 3]);
          }
          if (p.num_levels > 2) {  // 32 B per row, 32B swizzle (chunk ^= (row >> 2) & 1)
            float4* dst = reinterpret_cast<float4*>(set + EPI_L0 + EPI_L1 + trow * 32);
#pragma unroll
            for (int j = 0; j < 2; ++j)
              dst[j ^ ((trow >> 2) & 1)] = make_float4(l2[4 * j], l2[4 * j + 1], l2[4 * j + 2], l2[4 * j + 3]);
          }
          if (p.num_levels > 3) {  // 16 B per row, no swizzle
            *reinterpret_cast<float4*>(set + EPI_L0 + EPI_L1 + EPI_L2 + trow * 16) = make_float4(l3[0], l3[1], l3[2], l3[3]);
          }
          fence_proxy_async();
          epi_bar_sync();
          if (et == 0) {
            const int col = n0 + ch * CHUNK;
            const int mrow = m0 + mt * TILE_M;
            const uint32_t s0 = smem_u32(set);
            tma_store_3d(&map_l0, s0, col, mrow, row);
            if (p.num_levels > 1 && (col >> 1) < (p.W2 >> 1)) tma_store_3d(&map_l1, s0 + EPI_L0, col >> 1, mrow, row);
            if (p.num_levels > 2 && (col >> 2) < (p.W2 >> 2)) tma_store_3d(&map_l2, s0 + EPI_L0 + EPI_L1, col >> 2, mrow, row);
            if (p.num_levels > 3 && (col >> 3) < (p.W2 >> 3))
              tma_store_3d(&map_l3, s0 + EPI_L0 + EPI_L1 + EPI_L2, col >> 3, mrow, row);
            tma_store_commit();
          }
          chunk_parity ^= 1;
        }
      }
      acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
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
                    const cuuint32_t* box, CUtensorMapSwizzle swizzle, const char* what) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("corr1d_build(tf32): cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return NND_ERR_CUDA;
  }
  cuuint32_t elem_strides[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims,
                  strides_bytes, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("corr1d_build(tf32): cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, static_cast<int>(r));
    return NND_ERR_CUDA;
  }
  return NND_OK;
}

}  // namespace

nnd_status corr1d_build_tf32(const float* fmap1, const float* fmap2, int B, int C, int H, int W1, int W2,
                             int num_levels, float* const* level, const int* pitch, cudaStream_t stream) {
  // TMA constraints: every global stride a multiple of 16 bytes, bases 16-byte aligned
  if (W1 % 4 != 0 || W2 % 4 != 0 || !aligned16(fmap1) || !aligned16(fmap2)) {
    set_error("corr1d_build(tf32): TMA needs W1 and W2 to be multiples of 4 and 16-byte aligned feature maps "
              "(W1=%d, W2=%d); use NND_PREC_FP32 for this shape", W1, W2);
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
  p.m_groups = (W1 + 2 * TILE_M - 1) / (2 * TILE_M);
  p.n_chunks = (W2 + MAX_N - 1) / MAX_N;
  p.jobs = static_cast<long long>(p.rows) * p.m_groups * p.n_chunks;
  p.scale_div = static_cast<float>(sqrt(static_cast<double>(C)));
  p.scale_inv = 1.0f / p.scale_div;
  {
    int e = 0;
    p.scale_is_pow2 = (frexpf(p.scale_div, &e) == 0.5f) ? 1 : 0;
  }
  const int a_boxes = (min(W1, 2 * TILE_M) + BOX_W - 1) / BOX_W;
  const int b_boxes = (min(W2, MAX_N) + BOX_W - 1) / BOX_W;
  p.a_boxes_pad = (a_boxes + 3) & ~3;
  p.stage_bytes = (p.a_boxes_pad + b_boxes) * BOX_BYTES;
  const int budget = 227 * 1024 - 1024 /*alignment slack*/ - 2 * EPI_SET - 256 /*barriers*/;
  p.stages = budget / p.stage_bytes;
  if (p.stages > 6) p.stages = 6;
  if (p.stages < 2) {
    set_error("corr1d_build(tf32): a pipeline stage of %d bytes leaves no room for double buffering", p.stage_bytes);
    return NND_ERR_UNSUPPORTED;
  }
  const size_t smem_bytes = 1024 + static_cast<size_t>(p.stages) * p.stage_bytes + 2 * EPI_SET + 256;

  alignas(64) CUtensorMap map_a, map_b, map_l[4];
  {
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
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(CHUNK >> l), TILE_M, 1};
    nnd_status st = make_map(&map_l[l], level[ll], 3, dims, str, box, sw[l], "pyramid level");
    if (st != NND_OK) return st;
  }

  {
    cudaError_t e = cudaFuncSetAttribute(corr1d_build_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bytes));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(corr1d_build_tf32_kernel)");
  }
  const long long sms = sm_count();
  const unsigned grid = static_cast<unsigned>(p.jobs < sms ? p.jobs : sms);
  corr1d_build_tf32_kernel<<<grid, NUM_THREADS, smem_bytes, stream>>>(map_a, map_b, map_l[0], map_l[1], map_l[2],
                                                                      map_l[3], p);
  return check_launch("corr1d_build_tf32_kernel");
}

}  // namespace nnd
