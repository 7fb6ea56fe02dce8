// Correlation-volume builds with the pyramid pooled in the epilogue (sm_100a).
//
//   nnd_corr1d_build (fp32 path)  CorrBlock1D.corr + __init__      raft_stereo/cost_volume.py:55-61, :12-34
//   nnd_groupcorr_build           GroupCorrBlock1D.corr            raft_stereo/cost_volume.py:113-128
//                                 GeometryAwareCostVolume.build_cost_volume igev_stereo/cost_volume.py:81-98
//   nnd_avgpool_pairs             F.avg_pool1d(., 2)               raft_stereo/cost_volume.py:33
//
// The reference writes the volume with a batched SGEMM, re-reads it for the `/ C**0.5` pass and then
// once more per avg_pool1d level.  Here every output element is produced once in registers, scaled,
// and levels 0..3 are written straight from the accumulators: each thread owns 4 consecutive w2
// columns (levels 1 and 2 are in-thread sums, level 3 takes one shuffle with the neighbouring lane).
// Pooling follows avg_pool1d exactly -- level l+1 is (x[2j] + x[2j+1]) * 0.5 of the *rounded* level l
// values, odd tails dropped.
//
// The fp32 path is the 1e-5 parity mode (plain FFMA, sequential-in-c accumulation per output).  The
// tensor-core (tcgen05, TF32 operands) path lives in build_tcgen05.cu.
#include "common.cuh"

namespace nnd {

nnd_status corr1d_build_tf32(const float* fmap1, const float* fmap2, int B, int C, int H, int W1, int W2,
                             int num_levels, float* const* level, const int* pitch, cudaStream_t stream);
nnd_status corr1d_build_f16_nhwc(const void* fmap1, const void* fmap2, int B, int C, int H, int W1, int W2,
                                 int num_levels, float* const* level, const int* pitch, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// fp32 all-pairs build: per (b,h) row, out[m][n] = sum_c f1[c][m] * f2[c][n] / scale_div.
// 64x64 output tile per 128-thread block, 8(m) x 4(n) register tile per thread, BK = 16.
// Both operands are "k-slow" in NCHW (consecutive w for a fixed channel), which is exactly the
// [k][m] / [k][n] shared-memory layout an outer-product FFMA kernel wants: no transposes.
// ------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(128)
corr1d_build_fp32_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int C, int H, int W1, int W2,
                         float scale_div, int num_levels, Pyramid pyr, int vec_ok) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15;   // n direction: columns 4*tx .. 4*tx+3
  const int ty = tid >> 4;   // m direction: rows   8*ty .. 8*ty+7
  const int n_tile = blockIdx.x * BN;
  const int m_tile = blockIdx.y * BM;
  const long long bh = blockIdx.z;  // b*H + h
  const long long b = bh / H;
  const int h = static_cast<int>(bh - b * H);
  const long long plane1 = static_cast<long long>(H) * W1;
  const long long plane2 = static_cast<long long>(H) * W2;
  const float* a_row = f1 + b * C * plane1 + static_cast<long long>(h) * W1;  // + c*plane1 + m
  const float* b_row = f2 + b * C * plane2 + static_cast<long long>(h) * W2;  // + c*plane2 + n

  // loader role: 16 k-rows x 64 columns per operand = 1024 floats / 128 threads = 8 each
  const int lcol = tid & 63;
  const int lk0 = tid >> 6;  // 0..1, rows lk0, lk0+2, ...
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  auto fetch = [&](int c0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = lk0 + 2 * i;
      const int c = c0 + k;
      const int m = m_tile + lcol, n = n_tile + lcol;
      ra[i] = (c < C && m < W1) ? __ldg(a_row + c * plane1 + m) : 0.f;
      rb[i] = (c < C && n < W2) ? __ldg(b_row + c * plane2 + n) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      As[buf][lk0 + 2 * i][lcol] = ra[i];
      Bs[buf][lk0 + 2 * i][lcol] = rb[i];
    }
  };

  const int k_tiles = (C + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < k_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < k_tiles) fetch((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][8 * ty]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][8 * ty + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tx]);
      const float am[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
    }
    if (kt + 1 < k_tiles) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  const int n0 = n_tile + 4 * tx;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m_tile + 8 * ty + i;
    // every lane runs the shuffle inside store_row_quad; rows/columns out of range only skip stores
    float4 v;
    v.x = __fdiv_rn(acc[i][0], scale_div);
    v.y = __fdiv_rn(acc[i][1], scale_div);
    v.z = __fdiv_rn(acc[i][2], scale_div);
    v.w = __fdiv_rn(acc[i][3], scale_div);
    const long long row = bh * W1 + min(m, W1 - 1);
    store_row_quad(pyr, num_levels, row, n0, v, vec_ok != 0, m < W1);
  }
}

// ------------------------------------------------------------------------------------------------
// Group-wise build with a tiny contraction (K = group_size, 8 for IGEV / 4 for GroupCorrBlock1D):
// 2 flop per output byte, i.e. purely write-bandwidth bound -- provided the instruction stream keeps
// up (at 22 B/clk/SM of stores the budget is ~60 instructions per 16 output bytes).  One block per
// (b, g, h): the two K x W operand strips are staged once in shared memory.  A warp is 4 row-lanes x
// 8 column-lanes and owns groups of 16 volume rows: a thread keeps a[4 rows][K] in registers for the
// whole group and sweeps the row in 32-column steps, so every B value is loaded once per four outputs
// and a warp store covers four full 128-byte row segments.  The pooled levels come out of the same
// registers (levels 1-2 in-thread, level 3 with one shuffle).
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int gc_rows_per_thread(int k) { return k <= 8 ? 4 : k <= 16 ? 2 : 1; }  // a[TR][K] must stay in registers

template <int K>
__global__ void __launch_bounds__(256)
groupcorr_build_kernel(const float* __restrict__ f1, const float* __restrict__ f2, int C, int G, int H, int W1,
                       int W2, float scale_div, float scale_rcp, int num_levels, Pyramid pyr, int vec_ok) {
  constexpr int GC_TR = gc_rows_per_thread(K);
  constexpr int GC_ROWS = 4 * GC_TR;  // rows per warp group
  extern __shared__ __align__(16) float smem[];
  const int W2p = (W2 + 3) & ~3;
  float* As = smem;             // [K][W1]
  float* Bs = smem + K * W1;    // [K][W2p], zero padded
  const long long bgh = blockIdx.x;  // (b*G + g)*H + h
  const int h = static_cast<int>(bgh % H);
  const long long bg = bgh / H;
  const int g = static_cast<int>(bg % G);
  const long long b = bg / G;
  const long long plane1 = static_cast<long long>(H) * W1;
  const long long plane2 = static_cast<long long>(H) * W2;
  const float* a_src = f1 + (b * C + static_cast<long long>(g) * K) * plane1 + static_cast<long long>(h) * W1;
  const float* b_src = f2 + (b * C + static_cast<long long>(g) * K) * plane2 + static_cast<long long>(h) * W2;
  // stage the two K x W strips: 16-byte loads, all of a thread's loads issued before the first store
  // (a dependent load->store loop here costs one memory latency per iteration and dominated the kernel)
  if (vec_ok != 0 && (W1 & 3) == 0 && (W2 & 3) == 0 && aligned16(f1) && aligned16(f2)) {
    const int qa = W1 >> 2, qb = W2 >> 2;           // float4 per row
    const int total = K * (qa + qb);
    constexpr int MAXQ = 8;
    for (int base = threadIdx.x; base < total; base += MAXQ * blockDim.x) {
      float4 v[MAXQ];
#pragma unroll
      for (int j = 0; j < MAXQ; ++j) {
        const int i = base + j * blockDim.x;
        if (i < total) {
          const bool is_a = i < K * qa;
          const int ii = is_a ? i : i - K * qa;
          const int q = is_a ? qa : qb;
          const int k = ii / q, c = ii - k * q;
          v[j] = ldg_f4((is_a ? a_src + k * plane1 : b_src + k * plane2) + 4 * c);
        }
      }
#pragma unroll
      for (int j = 0; j < MAXQ; ++j) {
        const int i = base + j * blockDim.x;
        if (i < total) {
          const bool is_a = i < K * qa;
          const int ii = is_a ? i : i - K * qa;
          *reinterpret_cast<float4*>((is_a ? As : Bs) + 4 * ii) = v[j];  // W2p == W2 here: rows stay dense
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < K * W1; i += blockDim.x) {
      const int k = i / W1, m = i - k * W1;
      As[i] = __ldg(a_src + k * plane1 + m);
    }
    for (int i = threadIdx.x; i < K * W2p; i += blockDim.x) {
      const int k = i / W2p, n = i - k * W2p;
      Bs[i] = n < W2 ? __ldg(b_src + k * plane2 + n) : 0.f;
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  const int rl = lane >> 3;   // row-lane 0..3
  const int cl = lane & 7;    // column-lane 0..7: columns 4*cl .. 4*cl+3 of a 32-column sweep
  const int n_groups = (W1 + GC_ROWS - 1) / GC_ROWS;
  const int n_sweeps = (W2p + 31) / 32;
  const bool fast = vec_ok != 0 && (W2 & 7) == 0 && num_levels <= 4;  // every level's quad lies fully inside its row

  for (int grp = warp; grp < n_groups; grp += n_warps) {
    float a[GC_TR][K];
    int row_m[GC_TR];
#pragma unroll
    for (int i = 0; i < GC_TR; ++i) {
      row_m[i] = grp * GC_ROWS + rl + 4 * i;
      const int mm = min(row_m[i], W1 - 1);
#pragma unroll
      for (int k = 0; k < K; ++k) a[i][k] = As[k * W1 + mm];
    }
    for (int sw = 0; sw < n_sweeps; ++sw) {
      const int n0 = sw * 32 + 4 * cl;
      float4 acc[GC_TR];
#pragma unroll
      for (int i = 0; i < GC_TR; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 < W2p) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float4 bv = *reinterpret_cast<const float4*>(Bs + k * W2p + n0);
#pragma unroll
          for (int i = 0; i < GC_TR; ++i) {
            acc[i].x = fmaf(a[i][k], bv.x, acc[i].x);
            acc[i].y = fmaf(a[i][k], bv.y, acc[i].y);
            acc[i].z = fmaf(a[i][k], bv.z, acc[i].z);
            acc[i].w = fmaf(a[i][k], bv.w, acc[i].w);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < GC_TR; ++i) {
        float4 v;
        v.x = div_rn_fast(acc[i].x, scale_div, scale_rcp);
        v.y = div_rn_fast(acc[i].y, scale_div, scale_rcp);
        v.z = div_rn_fast(acc[i].z, scale_div, scale_rcp);
        v.w = div_rn_fast(acc[i].w, scale_div, scale_rcp);
        const bool row_ok = row_m[i] < W1;
        const long long row = bgh * W1 + min(row_m[i], W1 - 1);
        if (fast) {
          // W2 % 8 == 0: a quad inside level 0 has its pooled pair / single / shuffle partner inside too
          const bool ok = row_ok && n0 < W2;
          if (ok) *reinterpret_cast<float4*>(pyr.ptr[0] + row * pyr.pitch[0] + n0) = v;
          if (num_levels > 1) {
            const float l1a = pool2(v.x, v.y), l1b = pool2(v.z, v.w);
            if (ok) *reinterpret_cast<float2*>(pyr.ptr[1] + row * pyr.pitch[1] + (n0 >> 1)) = make_float2(l1a, l1b);
            if (num_levels > 2) {
              const float l2 = pool2(l1a, l1b);
              if (ok) pyr.ptr[2][row * pyr.pitch[2] + (n0 >> 2)] = l2;
              if (num_levels > 3) {
                const float other = __shfl_xor_sync(0xffffffffu, l2, 1);
                if (ok && (cl & 1) == 0) pyr.ptr[3][row * pyr.pitch[3] + (n0 >> 3)] = pool2(l2, other);
              }
            }
          }
        } else {
          store_row_quad(pyr, num_levels, row, n0, v, vec_ok != 0, row_ok);
        }
      }
    }
  }
}

// avg_pool1d(x, 2) for pyramid levels beyond the fused four.
__global__ void avgpool_pairs_kernel(const float* __restrict__ src, int src_pitch, float* __restrict__ dst,
                                     int dst_width, int dst_pitch, long long rows) {
  const long long total = rows * dst_width;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dst_width;
    const int j = static_cast<int>(i - r * dst_width);
    const float* s = src + r * src_pitch + 2 * j;
    dst[r * dst_pitch + j] = pool2(__ldg(s), __ldg(s + 1));
  }
}

nnd_status fill_pyramid(Pyramid& pyr, int W2, int num_levels, float* const* level, const int* pitch,
                               bool& vec_ok, const char* who) {
  NND_REQUIRE(level && pitch, "%s: null level/pitch array", who);
  NND_REQUIRE(num_levels >= 1 && num_levels <= NND_MAX_LEVELS, "%s: num_levels %d outside [1, %d]", who, num_levels,
              NND_MAX_LEVELS);
  memset(&pyr, 0, sizeof(pyr));
  vec_ok = true;
  int w = W2;
  for (int l = 0; l < num_levels; ++l) {
    NND_REQUIRE(w >= 1, "%s: level %d would be empty (W2 = %d)", who, l, W2);
    NND_REQUIRE(level[l], "%s: level %d pointer is null", who, l);
    NND_REQUIRE(pitch[l] >= w, "%s: level %d pitch %d < width %d", who, l, pitch[l], w);
    pyr.ptr[l] = level[l];
    pyr.width[l] = w;
    pyr.pitch[l] = pitch[l];
    if (l < 4) vec_ok = vec_ok && (pitch[l] % 4 == 0) && aligned16(level[l]);
    w >>= 1;
  }
  return NND_OK;
}

// levels 4.. are pooled from their predecessor by a separate pass (the reference never reads them)
nnd_status pool_tail(const Pyramid& pyr, int num_levels, long long rows, cudaStream_t stream) {
  for (int l = 4; l < num_levels; ++l) {
    const long long total = rows * pyr.width[l];
    const long long want = (total + 255) / 256;
    const int blocks = static_cast<int>(want < 148LL * 16 ? want : 148LL * 16);
    avgpool_pairs_kernel<<<blocks, 256, 0, stream>>>(pyr.ptr[l - 1], pyr.pitch[l - 1], pyr.ptr[l], pyr.width[l],
                                                     pyr.pitch[l], rows);
    nnd_status st = check_launch("avgpool_pairs_kernel");
    if (st != NND_OK) return st;
  }
  return NND_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of the volume builds (training): the gradient of level 0, contracted with the OTHER feature map.
//   which = 0:  d_f1[b, c, h, i] = 1/scale * sum_j dV[(b, g, h, i), j] * f2[b, c, h, j]
//   which = 1:  d_f2[b, c, h, j] = 1/scale * sum_i dV[(b, g, h, i), j] * f1[b, c, h, i]
// for c in group g = c / group_size (CorrBlock1D: one group of C channels).  Volume rows are (b, g, h, i) with `pitch`
// floats each.  Plain fp32 FFMA GEMM tiles (64 channels x 64 positions, k-step 16) through shared memory: the forward's
// precision class in fp32 mode, and exact enough (1e-6) to serve the tensor-core forward's backward as well.
// ------------------------------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256)
volume_grad_kernel(const float* __restrict__ dvol, int pitch, const float* __restrict__ fother, int C, int H, int W1, int W2,
                   int G, int group_size, float inv_scale, float* __restrict__ dout) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 1];   // [k][channel]
  __shared__ float Bs[TK][TN + 1];   // [k][position]
  const int bgh = blockIdx.z;                 // (b * G + g) * H + h
  const int h = bgh % H, bg = bgh / H, g = bg % G, b = bg / G;
  const int c0 = blockIdx.y * TM;             // channel within the group
  const int n0 = blockIdx.x * TN;             // output position (i for WHICH 0, j for WHICH 1)
  const int Wout = WHICH == 0 ? W1 : W2, Wk = WHICH == 0 ? W2 : W1;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, 4 x 4 outputs each
  const long long fo_base = ((static_cast<long long>(b) * C + g * group_size) * H + h) * Wk;        // + c * H * Wk + k
  const long long dv_base = (static_cast<long long>(bg) * H + h) * W1 * static_cast<long long>(pitch);  // + i * pitch + j
  float acc[4][4] = {};
  for (int k0 = 0; k0 < Wk; k0 += TK) {
    // A tile: fother[c0 + m][k0 + k]  (k contiguous)
    for (int e = tid; e < TM * TK; e += 256) {
      const int m = e / TK, k = e % TK;
      const int c = c0 + m, kk = k0 + k;
      As[k][m] = (c < group_size && kk < Wk) ? __ldg(fother + fo_base + static_cast<long long>(c) * H * Wk + kk) : 0.f;
    }
    // B tile: WHICH 0: dV[i = n0 + n][j = k0 + k] (k contiguous);  WHICH 1: dV[i = k0 + k][j = n0 + n] (n contiguous)
    for (int e = tid; e < TN * TK; e += 256) {
      int n, k;
      if (WHICH == 0) { n = e / TK; k = e % TK; } else { k = e / TN; n = e % TN; }
      const int nn = n0 + n, kk = k0 + k;
      const int i = WHICH == 0 ? nn : kk, j = WHICH == 0 ? kk : nn;
      Bs[k][n] = (nn < Wout && kk < Wk) ? __ldg(dvol + dv_base + static_cast<long long>(i) * pitch + j) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { a[r] = As[k][ty * 4 + r]; bb[r] = Bs[k][tx * 4 + r]; }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(a[r], bb[q], acc[r][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty * 4 + r;
    if (c >= group_size) continue;
    float* orow = dout + ((static_cast<long long>(b) * C + g * group_size + c) * H + h) * Wout;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int nn = n0 + tx * 4 + q;
      if (nn < Wout) orow[nn] = acc[r][q] * inv_scale;
    }
  }
}


}  // namespace nnd

extern "C" {

nnd_status nnd_corr1d_build(const float* fmap1, const float* fmap2, int B, int C, int H, int W1, int W2,
                            int num_levels, int precision, float* const* level, const int* pitch,
                            nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NND_REQUIRE(fmap1 && fmap2, "corr1d_build: null feature map");
  NND_REQUIRE(B > 0 && C > 0 && H > 0 && W1 > 0 && W2 > 0, "corr1d_build: B, C, H, W1, W2 must be positive");
  NND_REQUIRE(static_cast<long long>(B) * H <= 65535 * 1024LL, "corr1d_build: too many epipolar rows");
  Pyramid pyr;
  bool vec_ok;
  nnd_status st = fill_pyramid(pyr, W2, num_levels, level, pitch, vec_ok, "corr1d_build");
  if (st != NND_OK) return st;
  if (precision == NND_PREC_TF32) {
    st = corr1d_build_tf32(fmap1, fmap2, B, C, H, W1, W2, num_levels < 4 ? num_levels : 4, level, pitch, stream);
    if (st != NND_OK) return st;
    return pool_tail(pyr, num_levels, static_cast<long long>(B) * H * W1, stream);
  }
  NND_REQUIRE(precision == NND_PREC_FP32, "corr1d_build: unknown precision %d", precision);
  const float scale_div = static_cast<float>(sqrt(static_cast<double>(C)));
  const long long bh = static_cast<long long>(B) * H;
  NND_REQUIRE(bh <= 65535, "corr1d_build: B*H = %lld exceeds the fp32 path's grid limit (65535)", bh);
  dim3 grid((W2 + BN - 1) / BN, (W1 + BM - 1) / BM, static_cast<unsigned>(bh));
  corr1d_build_fp32_kernel<<<grid, 128, 0, stream>>>(fmap1, fmap2, C, H, W1, W2, scale_div, num_levels, pyr,
                                                     vec_ok ? 1 : 0);
  st = check_launch("corr1d_build_fp32_kernel");
  if (st != NND_OK) return st;
  return pool_tail(pyr, num_levels, bh * W1, stream);
}

nnd_status nnd_corr1d_build_nhwc_f16(const void* fmap1, const void* fmap2, int B, int C, int H, int W1, int W2,
                                     int num_levels, float* const* level, const int* pitch, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NND_REQUIRE(fmap1 && fmap2, "corr1d_build_nhwc_f16: null feature map");
  NND_REQUIRE(B > 0 && C > 0 && H > 0 && W1 > 0 && W2 > 0, "corr1d_build_nhwc_f16: B, C, H, W1, W2 must be positive");
  NND_REQUIRE(static_cast<long long>(B) * H <= 65535 * 1024LL, "corr1d_build_nhwc_f16: too many epipolar rows");
  Pyramid pyr;
  bool vec_ok;
  nnd_status st = fill_pyramid(pyr, W2, num_levels, level, pitch, vec_ok, "corr1d_build_nhwc_f16");
  if (st != NND_OK) return st;
  st = corr1d_build_f16_nhwc(fmap1, fmap2, B, C, H, W1, W2, num_levels < 4 ? num_levels : 4, level, pitch, stream);
  if (st != NND_OK) return st;
  return pool_tail(pyr, num_levels, static_cast<long long>(B) * H * W1, stream);
}

nnd_status nnd_groupcorr_build(const float* fmap1, const float* fmap2, int B, int C, int H, int W1, int W2,
                               int num_groups, int group_size, float scale_div, int num_levels,
                               float* const* level, const int* pitch, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NND_REQUIRE(fmap1 && fmap2, "groupcorr_build: null feature map");
  NND_REQUIRE(B > 0 && C > 0 && H > 0 && W1 > 0 && W2 > 0, "groupcorr_build: B, C, H, W1, W2 must be positive");
  NND_REQUIRE(num_groups > 0 && group_size > 0, "groupcorr_build: num_groups and group_size must be positive");
  // the reference indexes chunk i (of size group_size) for i < num_groups: IndexError beyond C
  NND_REQUIRE(static_cast<long long>(num_groups) * group_size <= C,
              "groupcorr_build: num_groups * group_size = %d exceeds C = %d", num_groups * group_size, C);
  NND_REQUIRE(scale_div > 0.f, "groupcorr_build: scale_div must be positive");
  Pyramid pyr;
  bool vec_ok;
  nnd_status st = fill_pyramid(pyr, W2, num_levels, level, pitch, vec_ok, "groupcorr_build");
  if (st != NND_OK) return st;
  const long long blocks = static_cast<long long>(B) * num_groups * H;
  NND_REQUIRE(blocks <= 0x7fffffffLL, "groupcorr_build: too many rows");
  const int W2p = (W2 + 3) & ~3;
  const size_t smem = static_cast<size_t>(group_size) * (W1 + W2p) * sizeof(float);
  NND_REQUIRE(smem <= 200 * 1024, "groupcorr_build: strips of %zu bytes do not fit shared memory", smem);
  // warps per block: the count in [4, 8] that divides the 16-row groups of a volume slab most evenly
  int gc_warps = 8;
  {
    const int group_rows = 4 * gc_rows_per_thread(group_size);
    const int n_groups = (W1 + group_rows - 1) / group_rows;
    double best = -1.0;
    for (int w = 8; w >= 4; --w) {
      const int rounds = (n_groups + w - 1) / w;
      const double eff = static_cast<double>(n_groups) / (rounds * w);
      if (eff > best + 1e-9) { best = eff; gc_warps = w; }
    }
  }
#define NND_LAUNCH_GROUP(KK)                                                                                      \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      cudaFuncSetAttribute(groupcorr_build_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                           static_cast<int>(smem));                                                               \
    groupcorr_build_kernel<KK><<<static_cast<unsigned>(blocks), 32 * gc_warps, smem, stream>>>(                   \
        fmap1, fmap2, C, num_groups, H, W1, W2, scale_div, 1.0f / scale_div, num_levels, pyr, vec_ok ? 1 : 0);    \
  } while (0)
  // every group size the reference's split quirk can produce with C <= 256 (G * G <= C: 1..16), and 32
  switch (group_size) {
    case 1: NND_LAUNCH_GROUP(1); break;
    case 2: NND_LAUNCH_GROUP(2); break;
    case 3: NND_LAUNCH_GROUP(3); break;
    case 4: NND_LAUNCH_GROUP(4); break;
    case 5: NND_LAUNCH_GROUP(5); break;
    case 6: NND_LAUNCH_GROUP(6); break;
    case 7: NND_LAUNCH_GROUP(7); break;
    case 8: NND_LAUNCH_GROUP(8); break;
    case 9: NND_LAUNCH_GROUP(9); break;
    case 10: NND_LAUNCH_GROUP(10); break;
    case 11: NND_LAUNCH_GROUP(11); break;
    case 12: NND_LAUNCH_GROUP(12); break;
    case 13: NND_LAUNCH_GROUP(13); break;
    case 14: NND_LAUNCH_GROUP(14); break;
    case 15: NND_LAUNCH_GROUP(15); break;
    case 16: NND_LAUNCH_GROUP(16); break;
    case 32: NND_LAUNCH_GROUP(32); break;
    default:
      set_error("groupcorr_build: group_size %d unsupported (1..16, 32)", group_size);
      return NND_ERR_UNSUPPORTED;
  }
#undef NND_LAUNCH_GROUP
  st = check_launch("groupcorr_build_kernel");
  if (st != NND_OK) return st;
  return pool_tail(pyr, num_levels, blocks * W1, stream);
}

nnd_status nnd_avgpool_pairs(const float* src, int src_width, int src_pitch, float* dst, int dst_pitch,
                             int64_t rows, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(src && dst, "avgpool_pairs: null pointer");
  NND_REQUIRE(src_width >= 2 && rows > 0, "avgpool_pairs: needs src_width >= 2 and rows > 0");
  const int dst_width = src_width / 2;
  NND_REQUIRE(src_pitch >= src_width && dst_pitch >= dst_width, "avgpool_pairs: pitch smaller than width");
  const long long total = rows * dst_width;
  const long long want = (total + 255) / 256;
  const int blocks = static_cast<int>(want < 148LL * 16 ? want : 148LL * 16);
  avgpool_pairs_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, src_pitch, dst, dst_width,
                                                                                   dst_pitch, rows);
  return check_launch("avgpool_pairs_kernel");
}

nnd_status nnd_volume_grad(const float* d_level0, int pitch, const float* f_other, int B, int C, int H, int W1, int W2,
                           int num_groups, int group_size, float scale_div, int which, float* d_fmap, nnd_stream_t stream) {
  using namespace nnd;
  NND_REQUIRE(d_level0 && f_other && d_fmap, "volume_grad: null pointer argument");
  NND_REQUIRE(B > 0 && C > 0 && H > 0 && W1 > 0 && W2 > 0, "volume_grad: B, C, H, W1, W2 must be positive");
  NND_REQUIRE(num_groups > 0 && group_size > 0 && static_cast<long long>(num_groups) * group_size <= C,
              "volume_grad: num_groups * group_size exceeds C");
  NND_REQUIRE(pitch >= W2, "volume_grad: pitch %d smaller than W2 %d", pitch, W2);
  NND_REQUIRE(scale_div > 0.f && (which == 0 || which == 1), "volume_grad: scale_div must be positive, which 0 or 1");
  const long long z = static_cast<long long>(B) * num_groups * H;
  NND_REQUIRE(z <= 65535, "volume_grad: B * G * H = %lld exceeds the grid limit (65535)", z);
  const int Wout = which == 0 ? W1 : W2;
  dim3 grid((Wout + 63) / 64, (group_size + 63) / 64, static_cast<unsigned>(z));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // channels beyond num_groups * group_size receive no gradient: the caller zero-fills d_fmap when they exist
  if (which == 0)
    volume_grad_kernel<0><<<grid, 256, 0, st>>>(d_level0, pitch, f_other, C, H, W1, W2, num_groups, group_size, 1.0f / scale_div, d_fmap);
  else
    volume_grad_kernel<1><<<grid, 256, 0, st>>>(d_level0, pitch, f_other, C, H, W1, W2, num_groups, group_size, 1.0f / scale_div, d_fmap);
  return check_launch("volume_grad_kernel");
}

}  // extern "C"
