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
// Interleaved ("g innermost") IGEV pyramids.
//
// In the reference layout ([b][g][h][w1] rows of D floats) the radius-4 window of one pixel is 8 groups
// x 40 bytes scattered over 8 rows: with 32-byte sectors and 64-byte DRAM granules the dual lookup
// drags 2.66 GB through HBM per iteration for 0.79 GB of useful window data (ncu, BASELINE config 4).
// The lookup-side pyramids are therefore stored pixel-major with the group index innermost:
//     level l:  [b][h][w1][d_l][g]   (row of one pixel = (D >> l) * G floats, 32-byte aligned)
// so the 8 windows of a pixel are ONE contiguous run of <= 11 * 32 bytes.  The layout is private to the
// cost-volume object; the reference-layout views (feat_corr_cv / geo_aware_cv) are synthesised on demand.
//
// Both producers below read the volume once and write levels 0..3 from registers.  A thread owns
// (pixel(s), 4 consecutive d, 4 consecutive g): levels 1 and 2 pool inside the thread, level 3 takes one
// shuffle with the thread holding the neighbouring d-quad (lane ^ 2).  Lane bit 0 selects the g-half, so
// the two lanes of a pair complete every 32-byte sector they touch in the same store instruction.
// Requires G == 8 and D % 8 == 0 (num_levels <= 4).
// ------------------------------------------------------------------------------------------------
struct InterleavedLevels {
  float* ptr[4];
  int num_levels;
};

// v[gi] = 4 consecutive d of group 4*gh + gi for ONE pixel; writes this thread's share of levels 0..3.
__device__ __forceinline__ void store_interleaved_quad(const InterleavedLevels& lv, long long pixrow, int D, int dq,
                                                       int gh, const float (&v)[4][4], bool ok) {
  // level 0: d = 4*dq + j  ->  floats (pixrow*D + d) * 8 + 4*gh
  if (ok) {
    float* p0 = lv.ptr[0] + (pixrow * D + 4 * dq) * 8 + 4 * gh;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(p0 + 8 * j) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
  }
  if (lv.num_levels < 2) return;
  float l1[4][2];
#pragma unroll
  for (int gi = 0; gi < 4; ++gi) {
    l1[gi][0] = pool2(v[gi][0], v[gi][1]);
    l1[gi][1] = pool2(v[gi][2], v[gi][3]);
  }
  if (ok) {
    float* p1 = lv.ptr[1] + (pixrow * (D >> 1) + 2 * dq) * 8 + 4 * gh;
    *reinterpret_cast<float4*>(p1) = make_float4(l1[0][0], l1[1][0], l1[2][0], l1[3][0]);
    *reinterpret_cast<float4*>(p1 + 8) = make_float4(l1[0][1], l1[1][1], l1[2][1], l1[3][1]);
  }
  if (lv.num_levels < 3) return;
  float l2[4];
#pragma unroll
  for (int gi = 0; gi < 4; ++gi) l2[gi] = pool2(l1[gi][0], l1[gi][1]);
  if (ok) *reinterpret_cast<float4*>(lv.ptr[2] + (pixrow * (D >> 2) + dq) * 8 + 4 * gh) = make_float4(l2[0], l2[1], l2[2], l2[3]);
  if (lv.num_levels < 4) return;
  float o[4];
#pragma unroll
  for (int gi = 0; gi < 4; ++gi) o[gi] = __shfl_xor_sync(0xffffffffu, l2[gi], 2);  // the d-quad dq ^ 1
  if (ok && (dq & 1) == 0)
    *reinterpret_cast<float4*>(lv.ptr[3] + (pixrow * (D >> 3) + (dq >> 1)) * 8 + 4 * gh) =
        make_float4(pool2(l2[0], o[0]), pool2(l2[1], o[1]), pool2(l2[2], o[2]), pool2(l2[3], o[3]));
}

// source = reference-layout volume: rows [b][g][h][w1] of D floats (d contiguous, row pitch `pitch`).
// lane = gh | (item_in_warp << 1); items = (pixel, d-quad) pairs, d-quad fastest.
__global__ void __launch_bounds__(256)
gev_interleave_dmajor_kernel(const float* __restrict__ vol, long long pitch, int D, long long hw, long long n_items,
                             InterleavedLevels lv) {
  const int gh = threadIdx.x & 1;
  const long long item = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 1;
  const bool ok = item < n_items;
  const int dqs = D >> 2;
  const long long it = ok ? item : 0;
  const long long pixrow = it / dqs;            // (b*H + h)*W1 + w1
  const int dq = static_cast<int>(it - pixrow * dqs);
  const long long b = pixrow / hw;
  const long long p = pixrow - b * hw;
  float v[4][4];
#pragma unroll
  for (int gi = 0; gi < 4; ++gi) {
    const long long row = (b * 8 + 4 * gh + gi) * hw + p;
    const float4 t = ok ? __ldcs(reinterpret_cast<const float4*>(vol + row * pitch + 4 * dq)) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[gi][0] = t.x; v[gi][1] = t.y; v[gi][2] = t.z; v[gi][3] = t.w;
  }
  store_interleaved_quad(lv, pixrow, D, dq, gh, v, ok);
}

// source = regulariser output (B, 8, D, H, W1), w1 contiguous.
// lane = gh | (dq parity << 1) | (w-quad in a group of 8 << 2); warp item = (b, h, w-group of 32 px, d-octet).
__global__ void __launch_bounds__(256)
gev_interleave_wmajor_kernel(const float* __restrict__ geo, int D, int H, int W1, int w_groups, long long n_warp_items,
                             InterleavedLevels lv) {
  const int lane = threadIdx.x & 31;
  const long long witem = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (witem >= n_warp_items) return;  // warp-uniform
  const int gh = lane & 1, dqp = (lane >> 1) & 1, wql = lane >> 2;
  const int d_octs = D >> 3;
  long long t = witem;
  const int doct = static_cast<int>(t % d_octs);
  t /= d_octs;
  const int wg = static_cast<int>(t % w_groups);
  const long long bh = t / w_groups;
  const long long b = bh / H;
  const int h = static_cast<int>(bh - b * H);
  const int w0 = (wg * 8 + wql) * 4;
  const bool ok = w0 < W1;  // W1 % 4 == 0: a quad is inside or outside as a whole
  const int dq = doct * 2 + dqp;
  const long long plane = static_cast<long long>(H) * W1;
  float4 raw[4][4];  // [gi][j]: 4 pixels of (g = 4*gh + gi, d = 4*dq + j)
#pragma unroll
  for (int gi = 0; gi < 4; ++gi)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      raw[gi][j] = ok ? __ldcs(reinterpret_cast<const float4*>(
                            geo + ((b * 8 + 4 * gh + gi) * D + 4 * dq + j) * plane + static_cast<long long>(h) * W1 + w0))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v[4][4];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[gi][j] = e == 0 ? raw[gi][j].x : e == 1 ? raw[gi][j].y : e == 2 ? raw[gi][j].z : raw[gi][j].w;
    store_interleaved_quad(lv, bh * W1 + w0 + e, D, dq, gh, v, ok);
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
  // disparities slice, slice + S, slice + 2S, ... ; NF loads in flight per lane
#ifndef NND_SA_INFLIGHT
#define NND_SA_INFLIGHT 4      // loads in flight per lane and round; with four disparity slices per block at cfg4 this
#endif                        // measured 41.0 us (8 loads, one slice: 45.1; 16 loads: 61; 2 loads, 8 slices: 47)
  constexpr int NF = NND_SA_INFLIGHT;
  for (int d = slice; d < D; d += NF * S) {
    float v[NF][VEC];
    int n_valid = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) v[i][c] = 0.f;
      if (d + i * S < D) {
        SoftVec<VEC>::unpack(__ldcs(src + (d + i * S) * hwv), v[i]);
        n_valid = i + 1;
      }
    }
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      float zc[NF];
#pragma unroll
      for (int i = 0; i < NF; ++i) zc[i] = v[i][c];
      soft_push<NF>(st[c], zc, d, S, n_valid);
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

// slices per block: enough disparity slices for ~48 warps per SM (many warps with few loads each stream better here than
// few warps with deep unrolling); the merge cost grows with S, so at most 8.
static int soft_argmin_slices(long long units, int D) {
#ifndef NND_SA_WANT
#define NND_SA_WANT 48
#endif
  const long long want = static_cast<long long>(sm_count()) * NND_SA_WANT;
  int S = 1;
  while (S < 8 && units * S < want && 2 * S <= D) S *= 2;
  return S;
}

// ---------------------------------------------------------------------------------------------------
// cv_squeezer (Conv3d(8 -> 1, 3x3x3, padding 1), igev_stereo/model.py:65,144-145) fused with the soft-argmin
// (model.py:92-95,146) on the interleaved level-0 geometry volume [b][h][w][d][g].
//
// Thread = one input disparity plane d' with its 8 groups in registers (one 32-byte load per pixel, lanes run
// along d' = contiguous memory).  Instead of gathering, every thread SCATTERS its plane: along d into the three
// partial sums A_kd[d'] = sum_{g,kh,kw} W[g,kd,kh,kw] * x[g,d',..] (combined across lanes only once per output
// row, through shared memory: out[d] = bias + A_0[d-1] + A_1[d] + A_2[d+1], fixed order, deterministic), and
// along h into the three output rows an input row touches.  A warp-set (ceil(D/32) warps) marches down a
// strip of SQ_TW pixel columns: each input row is loaded ONCE (halo only in w: SQ_TW + 2 pixels), feeds
// 3 rows x SQ_TW pixels x 3 d-taps = 36 register accumulators with 864 FMAs (432 packed FFMA2), then the finished row's cost is
// assembled and one warp per pixel runs the softmax-expectation.  The 216 weights are uniform constant-bank
// operands (two LDCU.128 per tap serve SQ_TW pixels).  SQ_NS strips per CTA keep the warp count a multiple
// of four at D = 160 (20 warps) and share the halo pixels through L1.
// ---------------------------------------------------------------------------------------------------
constexpr int SQ_TW = 4;
// Conv3d weight (1, 8, 3, 3, 3) repacked [kd][kh][kw][g] (the eight group weights of a tap = two uniform
// 16-byte constant loads), bias in [54].x.  Filled stream-ordered from g_squeeze_stage by the ABI call.
__constant__ float4 c_squeeze[56];
__device__ float g_squeeze_stage[224];

__global__ void squeeze_pack_kernel(const float* __restrict__ weight, const float* __restrict__ bias) {
  const int i = threadIdx.x;  // destination index ((kd*3 + kh)*3 + kw)*8 + g
  if (i < 216) g_squeeze_stage[i] = weight[(i & 7) * 27 + (i >> 3)];
  if (i >= 216 && i < 224) g_squeeze_stage[i] = (i == 216 && bias) ? bias[0] : 0.f;
}

struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ldg_f8(const float* p) {  // 32-byte aligned, read-only path, one LDG.256
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {   // (a.x*b.x + c.x, a.y*b.y + c.y), one FFMA2
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// grid = (w-tiles, h-segments, B); block = n_strips * n_chunks warps; seg_rows output rows per CTA
__global__ void __launch_bounds__(640)
gev_squeeze_soft_argmin_kernel(const float* __restrict__ geo, int D, int H, int W, int n_chunks, int seg_rows,
                               float* __restrict__ out, float* __restrict__ cost_out) {
  extern __shared__ float sq_part[];  // [2 buffers][strip][3 kd][SQ_TW][D + 2], index d' + 1; [0] and [D + 1] stay zero
  const int Dp = D + 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strip = warp / n_chunks, chunk = warp - strip * n_chunks;
  const int n_strips = (blockDim.x >> 5) / n_chunks;
  const int dq = chunk * 32 + lane;
  const int w0 = (blockIdx.x * n_strips + strip) * SQ_TW;
  const int b = blockIdx.z;
  const int hs = blockIdx.y * seg_rows, he = min(H, hs + seg_rows);
  const bool live = dq < D && w0 < W;
  const int strip_floats = 3 * SQ_TW * Dp;
  float* my = sq_part + strip * strip_floats;           // buffer 0 of this strip; buffer 1 at + n_strips * strip_floats
  const int buf_stride = n_strips * strip_floats;

  for (int i = threadIdx.x; i < 2 * n_strips * 3 * SQ_TW; i += blockDim.x) {
    sq_part[i * Dp] = 0.f;
    sq_part[i * Dp + D + 1] = 0.f;
  }
  __syncthreads();

  // accumulators of the output rows h' - 1 (prev), h' (cur), h' + 1 (next) while input row h' is processed
  float acc[3][SQ_TW][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int p = 0; p < SQ_TW; ++p) acc[r][p][0] = acc[r][p][1] = acc[r][p][2] = 0.f;

  const float bias = c_squeeze[54].x;
  for (int hh = hs - 1; hh <= he; ++hh) {     // input rows; rows outside the image are zero padding
    if (live && hh >= 0 && hh < H) {
      const float* row = geo + (((static_cast<long long>(b) * H + hh) * W) * D + dq) * 8;
      F8 x[SQ_TW + 2];
#pragma unroll
      for (int nw = 0; nw < SQ_TW + 2; ++nw) {
        const int ww = w0 + nw - 1;
        if (ww >= 0 && ww < W) {
          x[nw] = ldg_f8(row + static_cast<long long>(ww) * D * 8);
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) x[nw].v[g] = 0.f;
        }
      }
      // input row hh reaches output row hh + 1 - kh: acc[2 - kh]  (kh = 0 -> next, 1 -> cur, 2 -> prev)
      // Packed FMAs (sm_100 FFMA2): the even and the odd groups of a plane accumulate in the two halves of one register
      // pair -- the x pair is two adjacent registers of the 32-byte load, the weight pair two adjacent words of the
      // constant bank (a uniform-register pair operand), so nothing has to be moved to form the pairs.  A (kh, kd, pixel)
      // chain runs over its three kw taps: 12 FFMA2 + 2 FADD instead of 24 FFMA (726 -> 656 us at cfg4).
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
          for (int ow = 0; ow < SQ_TW; ++ow) {
            float2 p = make_float2(0.f, 0.f);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float4 wa = c_squeeze[((kd * 3 + kh) * 3 + kw) * 2], wb = c_squeeze[((kd * 3 + kh) * 3 + kw) * 2 + 1];
              const F8& v = x[ow + kw];
              p = fma2(make_float2(v.v[0], v.v[1]), make_float2(wa.x, wa.y), p);
              p = fma2(make_float2(v.v[2], v.v[3]), make_float2(wa.z, wa.w), p);
              p = fma2(make_float2(v.v[4], v.v[5]), make_float2(wb.x, wb.y), p);
              p = fma2(make_float2(v.v[6], v.v[7]), make_float2(wb.z, wb.w), p);
            }
            acc[2 - kh][ow][kd] += p.x + p.y;
          }
        }
      }
    }
    // output row ho = hh - 1 is complete
    const int ho = hh - 1;
    float* buf = my + (ho & 1) * buf_stride;
    if (ho >= hs) {   // block-uniform
      if (live) {
#pragma unroll
        for (int p = 0; p < SQ_TW; ++p)
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) buf[(kd * SQ_TW + p) * Dp + dq + 1] = acc[0][p][kd];
      }
      // one barrier per row and per strip (the strips of a CTA are independent): the buffers alternate, row
      // ho + 2 is written after the next barrier
      asm volatile("bar.sync %0, %1;" ::"r"(strip + 1), "r"(n_chunks * 32) : "memory");
      for (int p = chunk; p < SQ_TW; p += n_chunks) {
        const int ww = w0 + p;
        if (ww >= W) continue;  // warp-uniform
        const float* a0 = buf + (0 * SQ_TW + p) * Dp;
        const float* a1 = buf + (1 * SQ_TW + p) * Dp;
        const float* a2 = buf + (2 * SQ_TW + p) * Dp;
        // out[d] = bias + A_0[d' = d - 1] + A_1[d' = d] + A_2[d' = d + 1]  (tap kd reads plane d + kd - 1, as Conv3d does)
        float mx = -FLT_MAX;
        for (int d = lane; d < D; d += 32) {
          const float c = ((bias + a0[d]) + a1[d + 1]) + a2[d + 2];
          if (cost_out) cost_out[((static_cast<long long>(b) * D + d) * H + ho) * W + ww] = c;
          mx = fmaxf(mx, c * LOG2E);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f, ws = 0.f;
        for (int d = lane; d < D; d += 32) {
          const float c = ((bias + a0[d]) + a1[d + 1]) + a2[d + 2];
          const float e = ex2_fast(fmaf(c, LOG2E, -mx));
          sum += e;
          ws = fmaf(static_cast<float>(d), e, ws);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          ws += __shfl_xor_sync(0xffffffffu, ws, o);
        }
        if (lane == 0) out[(static_cast<long long>(b) * H + ho) * W + ww] = -(ws / sum);
      }
    }
    // rotate: prev <- cur, cur <- next, next <- 0
#pragma unroll
    for (int p = 0; p < SQ_TW; ++p)
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        acc[0][p][kd] = acc[1][p][kd];
        acc[1][p][kd] = acc[2][p][kd];
        acc[2][p][kd] = 0.f;
      }
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

nnd_status nnd_gev_interleave_pool(const float* vol, int layout, int src_pitch, int B, int G, int D, int H, int W1,
                                   int num_levels, float* const* level, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NND_REQUIRE(vol && level, "gev_interleave_pool: null pointer");
  NND_REQUIRE(B > 0 && D > 0 && H > 0 && W1 > 0, "gev_interleave_pool: B, D, H, W1 must be positive");
  NND_REQUIRE(G == 8, "gev_interleave_pool: the interleaved layout is built for 8 groups (got %d)", G);
  NND_REQUIRE(D % 8 == 0, "gev_interleave_pool: D = %d must be a multiple of 8", D);
  NND_REQUIRE(num_levels >= 1 && num_levels <= 4, "gev_interleave_pool: num_levels %d outside [1, 4]", num_levels);
  NND_REQUIRE(layout == 0 || layout == 1, "gev_interleave_pool: layout %d is not 0 (rows of D) or 1 (B,G,D,H,W)", layout);
  NND_REQUIRE(aligned16(vol), "gev_interleave_pool: source must be 16-byte aligned");
  InterleavedLevels lv;
  memset(&lv, 0, sizeof(lv));
  lv.num_levels = num_levels;
  for (int l = 0; l < num_levels; ++l) {
    NND_REQUIRE(level[l] && aligned16(level[l]), "gev_interleave_pool: level %d pointer is null or unaligned", l);
    lv.ptr[l] = level[l];
  }
  const long long hw = static_cast<long long>(H) * W1;
  if (layout == 0) {
    NND_REQUIRE(src_pitch >= D && src_pitch % 4 == 0, "gev_interleave_pool: row pitch %d must be >= D and a multiple of 4",
                src_pitch);
    const long long n_items = static_cast<long long>(B) * hw * (D / 4);
    const long long blocks = (2 * n_items + 255) / 256;
    NND_REQUIRE(blocks <= 0x7fffffffLL, "gev_interleave_pool: volume too large for one launch");
    gev_interleave_dmajor_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(vol, src_pitch, D, hw, n_items, lv);
    return check_launch("gev_interleave_dmajor_kernel");
  }
  NND_REQUIRE(W1 % 4 == 0, "gev_interleave_pool: layout 1 needs W1 %% 4 == 0 (got %d)", W1);
  const int w_groups = (W1 + 31) / 32;
  const long long n_warp_items = static_cast<long long>(B) * H * w_groups * (D / 8);
  const long long blocks = (n_warp_items + 7) / 8;
  NND_REQUIRE(blocks <= 0x7fffffffLL, "gev_interleave_pool: volume too large for one launch");
  gev_interleave_wmajor_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(vol, D, H, W1, w_groups, n_warp_items, lv);
  return check_launch("gev_interleave_wmajor_kernel");
}

nnd_status nnd_gev_squeeze_soft_argmin(const float* geo_level0, const float* weight, const float* bias, int B, int G,
                                       int D, int H, int W1, float* out, float* cost_out, nnd_stream_t stream_) {
  using namespace nnd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  NND_REQUIRE(geo_level0 && weight && out, "gev_squeeze_soft_argmin: null pointer");
  NND_REQUIRE(G == 8, "gev_squeeze_soft_argmin: the interleaved layout holds 8 groups (got %d)", G);
  NND_REQUIRE(B > 0 && D > 0 && H > 0 && W1 > 0, "gev_squeeze_soft_argmin: B, D, H, W1 must be positive");
  NND_REQUIRE(B <= 65535, "gev_squeeze_soft_argmin: B exceeds the grid limit");
  NND_REQUIRE(D <= 640, "gev_squeeze_soft_argmin: D = %d exceeds 640 (one thread per disparity plane, 20 warps)", D);
  NND_REQUIRE((reinterpret_cast<uintptr_t>(geo_level0) & 31u) == 0, "gev_squeeze_soft_argmin: volume must be 32-byte aligned");
  // weights -> constant bank, stream-ordered (no host synchronisation; capturable).  Launches on different
  // streams with DIFFERENT weights would race on the symbol; one model per process uses one squeezer.
  squeeze_pack_kernel<<<1, 224, 0, stream>>>(weight, bias);
  nnd_status ps = check_launch("squeeze_pack_kernel");
  if (ps != NND_OK) return ps;
  void* stage = nullptr;
  cudaError_t e = cudaGetSymbolAddress(&stage, g_squeeze_stage);
  if (e != cudaSuccess) return cuda_fail(e, "gev_squeeze_soft_argmin: staging symbol");
  e = cudaMemcpyToSymbolAsync(c_squeeze, stage, 224 * sizeof(float), 0, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "gev_squeeze_soft_argmin: weight upload");
  const int n_chunks = (D + 31) / 32;
  const int w_strips = (W1 + SQ_TW - 1) / SQ_TW;
  int n_strips = 20 / n_chunks;                       // up to 20 warps per CTA (a multiple of four at D = 160)
  if (n_strips < 1) n_strips = 1;
  if (n_strips > w_strips) n_strips = w_strips;
  const int w_tiles = (w_strips + n_strips - 1) / n_strips;
  // h-segments: about four CTAs per SM in total; every segment re-reads two halo rows
  const long long cols = static_cast<long long>(B) * w_tiles;
  long long segs = (6LL * sm_count() + cols - 1) / cols;
  if (segs < 1) segs = 1;
  int seg_rows = static_cast<int>((H + segs - 1) / segs);
  if (seg_rows < 8) seg_rows = H < 8 ? H : 8;
  const int h_segs = (H + seg_rows - 1) / seg_rows;
  NND_REQUIRE(h_segs <= 65535, "gev_squeeze_soft_argmin: H = %d exceeds the grid limit", H);
  const int threads = n_strips * n_chunks * 32;
  const size_t smem = static_cast<size_t>(2) * n_strips * 3 * SQ_TW * (D + 2) * sizeof(float);
  NND_REQUIRE(smem <= 200 * 1024, "gev_squeeze_soft_argmin: D = %d needs %zu bytes of shared memory", D, smem);
  // per call (once per forward) rather than cached in a static: the attribute is per device
  e = cudaFuncSetAttribute(gev_squeeze_soft_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_fail(e, "gev_squeeze_soft_argmin: shared-memory attribute");
  dim3 grid(w_tiles, h_segs, B);
  gev_squeeze_soft_argmin_kernel<<<grid, threads, smem, stream>>>(geo_level0, D, H, W1, n_chunks, seg_rows, out, cost_out);
  return check_launch("gev_squeeze_soft_argmin_kernel");
}

}  // extern "C"
