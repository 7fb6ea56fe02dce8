// Fused pyramid lookups (sm_100a): every level, every tap, NCHW output, one launch.
//
//   nnd_corr1d_lookup          CorrBlock1D.__call__            raft_stereo/cost_volume.py:36-53
//   nnd_group_lookup  mode 0   GeometryAwareCostVolume.forward igev_stereo/cost_volume.py:54-79
//                     mode 1   GroupCorrBlock1D.__call__       raft_stereo/cost_volume.py:92-111
//   nnd_corr1d_lookup_indices  the integer half of linear_sampler, raft_stereo/utils.py:16-21
//
// Work decomposition: one warp = 32 consecutive pixels x one pyramid level (x a chunk of planes:
// plane = (source pyramid, group)).  A pixel's window at a level is <= 2r+3 consecutive floats of
// its own volume row, so neighbouring pixels never share data and a thread-per-pixel gather would
// cost one L1 wavefront per lane per tap.  Instead the warp loads the 32 windows cooperatively --
// WINQ lanes per pixel, one 16-byte load each, 32/WINQ rows per instruction -- parks them in a
// bank-swizzled shared-memory tile and then every lane interpolates its own pixel's taps from
// shared memory.  That is the minimum number of L1 wavefronts (one per touched row segment), the
// loads of a warp are all independent (WINQ in flight per lane), and the stores are 128-byte
// coalesced channel planes.
//
// Bit-exactness: tap positions follow the reference's fp32 operation order exactly
// (common.cuh:sampler_position); integer indices therefore equal the CPU reference's, and the lerp
// is evaluated without FMA contraction.
#include "common.cuh"

namespace nnd {

struct LookupArgs {
  ConstPyramid src[2];  // [0] = feature correlation, [1] = geometry volume (IGEV only)
  const float* coords;
  float* out;
  long long hw;     // H * W1
  long long n_pix;  // B * H * W1
  int G;            // planes (groups) per pixel and source
  int n_src;        // 1 or 2
  int num_levels;
  int radius;
  int mode;              // 0: channel = l*(S*G*T) + s*(G*T) + g*T + k ; 1: GroupCorrBlock1D view quirk
  int planes_per_block;  // chunk of the S*G planes handled by one block (blockIdx.y selects it)
  int vec;               // 1: pitches % 4 == 0 and bases 16-byte aligned -> float4 loads
};

// Window bookkeeping for one (pixel, level): [lo, hi] is the index range the taps can touch.
struct TapRange {
  int lo, hi;
};

__device__ __forceinline__ float tap_x(int k, int r, float centre) {
  // dx + coords / 2**i  (cost_volume.py:44-46): dx = k - r is an exact small integer.
  return __fadd_rn(static_cast<float>(k - r), centre);
}

__device__ __forceinline__ TapRange tap_range(float centre, int r, float span) {
  const float t_lo = sampler_position(tap_x(0, r, centre), span);
  const float t_hi = sampler_position(tap_x(2 * r, r, centre), span);
  TapRange tr;
  tr.lo = static_cast<int>(floorf(t_lo));
  tr.hi = static_cast<int>(ceilf(t_hi));
  return tr;
}

// WINQ : 16-byte quads per staged window (window = 4*WINQ floats >= 2r+6)
// TAPS : compile-time tap count (2r+1) -> per-tap state lives in registers across planes; 0 = dynamic
template <int WINQ, int TAPS>
__global__ void __launch_bounds__(32 * NND_MAX_LEVELS)
pyramid_lookup_kernel(const LookupArgs a) {
  constexpr int COLS = 4 * WINQ;
  constexpr int PPL = 32 / WINQ;  // pixels per cooperative load instruction
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ float smem[];

  const int lane = threadIdx.x;
  const int lvl = threadIdx.y;
  float* win = smem + lvl * (COLS * 32);

  const long long pix0 = static_cast<long long>(blockIdx.x) * 32;
  const long long pix = pix0 + lane;
  const bool valid = pix < a.n_pix;
  const int r = a.radius;
  const int T = TAPS > 0 ? TAPS : 2 * r + 1;
  const int w = a.src[0].width[lvl];
  const int pitch = a.src[0].pitch[lvl];
  const float span = static_cast<float>(w - 1);

  const float c = valid ? __ldg(a.coords + pix) : 0.0f;
  // coords / 2**lvl: scaling by a power of two is exact, so the product equals the IEEE quotient.
  const float centre = __fmul_rn(c, 1.0f / static_cast<float>(1 << lvl));
  const TapRange tr = tap_range(centre, r, span);
  const int s = tr.lo & ~3;  // window start, quad aligned (rows are 16-byte aligned in vec mode)

  // per-tap state: shared-memory word offsets of the two neighbours and the lerp weights
  int o0[TAPS > 0 ? TAPS : 1], o1[TAPS > 0 ? TAPS : 1];
  float cf[TAPS > 0 ? TAPS : 1], omc[TAPS > 0 ? TAPS : 1];
  auto tap_setup = [&](int k, int& off0, int& off1, float& coef, float& one_minus) {
    const float t = sampler_position(tap_x(k, r, centre), span);
    const float f0 = floorf(t), f1 = ceilf(t);
    int i0 = static_cast<int>(f0) - s, i1 = static_cast<int>(f1) - s;
    i0 = min(max(i0, 0), COLS - 1);
    i1 = min(max(i1, 0), COLS - 1);
    off0 = i0 * 32 + ((lane + PPL * (i0 >> 2)) & 31);
    off1 = i1 * 32 + ((lane + PPL * (i1 >> 2)) & 31);
    coef = __fsub_rn(f1, t);               // coef = idx1 - t            (utils.py:26)
    one_minus = __fsub_rn(1.0f, coef);     // (1 - coef)                 (utils.py:27)
  };
  if (TAPS > 0) {
#pragma unroll
    for (int k = 0; k < TAPS; ++k) tap_setup(k, o0[k], o1[k], cf[k], omc[k]);
  }

  // cooperative loader role: this lane fetches quad q of pixel (j*PPL + lane/WINQ), j < WINQ
  const int q = lane % WINQ;
  long long row0[WINQ];  // volume row of (pixel, group 0)
  int col0[WINQ];        // first column of my quad, or -1 when the quad is not needed
  int slot[WINQ];        // swizzled position of that pixel inside a column of the tile
#pragma unroll
  for (int j = 0; j < WINQ; ++j) {
    const int p = j * PPL + lane / WINQ;
    const int sp = __shfl_sync(FULL, s, p);
    const int hp = __shfl_sync(FULL, tr.hi, p);
    const long long ppix = pix0 + p;
    const int cq = sp + 4 * q;
    const bool need = ppix < a.n_pix && cq <= hp && cq < w;
    const long long pb = ppix / a.hw;
    row0[j] = pb * a.G * a.hw + (ppix - pb * a.hw);
    col0[j] = need ? cq : -1;
    slot[j] = (p + PPL * q) & 31;
  }

  const long long b = pix / a.hw;
  const long long rem = pix - b * a.hw;
  const int GT = a.G * T;
  const long long c_total = static_cast<long long>(a.num_levels) * a.n_src * GT;
  const int n_planes = a.n_src * a.G;
  const int plane_begin = blockIdx.y * a.planes_per_block;
  const int plane_end = min(plane_begin + a.planes_per_block, n_planes);

  for (int plane = plane_begin; plane < plane_end; ++plane) {
    const int sidx = plane / a.G;
    const int g = plane - sidx * a.G;
    const float* __restrict__ base = a.src[sidx].ptr[lvl];

    float4 v[WINQ];
#pragma unroll
    for (int j = 0; j < WINQ; ++j) {
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col0[j] >= 0) {
        const float* p = base + (row0[j] + g * a.hw) * pitch + col0[j];
        if (a.vec) {
          v[j] = ldg_f4(p);
        } else {
          const int left = w - col0[j];
          v[j].x = __ldg(p);
          if (left > 1) v[j].y = __ldg(p + 1);
          if (left > 2) v[j].z = __ldg(p + 2);
          if (left > 3) v[j].w = __ldg(p + 3);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < WINQ; ++j) {
      float* dst = win + (4 * q) * 32 + slot[j];
      dst[0] = v[j].x;
      dst[32] = v[j].y;
      dst[64] = v[j].z;
      dst[96] = v[j].w;
    }
    __syncwarp();

    if (valid) {
      long long out_base;
      if (a.mode == 0) {
        const long long ch0 = static_cast<long long>(lvl) * a.n_src * GT + static_cast<long long>(sidx) * GT + g * T;
        out_base = (b * c_total + ch0) * a.hw + rem;
      } else {
        out_base = (b * c_total + static_cast<long long>(lvl) * GT) * a.hw;
      }
      for (int k = 0; k < T; ++k) {
        int off0, off1;
        float coef, one_minus;
        if (TAPS > 0) {
          off0 = o0[k]; off1 = o1[k]; coef = cf[k]; one_minus = omc[k];
        } else {
          tap_setup(k, off0, off1, coef, one_minus);
        }
        const float v0 = win[off0];
        const float v1 = win[off1];
        // coef * val0 + (1 - coef) * val1, each operation rounded (utils.py:27)
        const float res = __fadd_rn(__fmul_rn(coef, v0), __fmul_rn(one_minus, v1));
        if (a.mode == 0) {
          a.out[out_base + static_cast<long long>(k) * a.hw] = res;
        } else {
          // reference memory order is [b][g][h][w][k]; it is *viewed* as (B, H, W, G*T)
          // (raft_stereo/cost_volume.py:108) and then permuted to NCHW.
          const long long f = (static_cast<long long>(g) * a.hw + rem) * T + k;
          const long long cprime = f % GT;
          const long long pos = f / GT;
          a.out[out_base + cprime * a.hw + pos] = res;
        }
      }
    }
    __syncwarp();
  }
}

__global__ void lookup_indices_kernel(const int* __restrict__ width_dev, const float* __restrict__ coords,
                                      long long n_pix, int num_levels, int radius, int w0, int w1, int w2, int w3,
                                      int w4, int w5, int w6, int w7, int32_t* __restrict__ idx0,
                                      int32_t* __restrict__ idx1) {
  (void)width_dev;
  const int T = 2 * radius + 1;
  const long long total = n_pix * T * num_levels;
  const int widths[NND_MAX_LEVELS] = {w0, w1, w2, w3, w4, w5, w6, w7};
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % T);
    const long long pl = i / T;
    const long long pix = pl % n_pix;
    const int lvl = static_cast<int>(pl / n_pix);
    const float span = static_cast<float>(widths[lvl] - 1);
    const float centre = __fmul_rn(__ldg(coords + pix), 1.0f / static_cast<float>(1 << lvl));
    const float t = sampler_position(tap_x(k, radius, centre), span);
    idx0[i] = static_cast<int32_t>(floorf(t));
    idx1[i] = static_cast<int32_t>(ceilf(t));
  }
}

static nnd_status launch_lookup(const float* const* level_a, const float* const* level_b, const int* width,
                                const int* pitch, const float* coords, int B, int G, int H, int W1, int num_levels,
                                int radius, int n_src, int mode, float* out, cudaStream_t stream) {
  NND_REQUIRE(level_a && width && pitch && coords && out, "lookup: null pointer argument");
  NND_REQUIRE(B > 0 && G > 0 && H > 0 && W1 > 0, "lookup: B, G, H, W1 must be positive (got %d %d %d %d)", B, G, H, W1);
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
  a.hw = static_cast<long long>(H) * W1;
  a.n_pix = a.hw * B;
  a.G = G;
  a.n_src = n_src;
  a.num_levels = num_levels;
  a.radius = radius;
  a.mode = mode;
  a.vec = vec ? 1 : 0;
  const int n_planes = n_src * G;
  a.planes_per_block = n_planes >= 8 ? 4 : n_planes;

  const long long blocks_x = (a.n_pix + 31) / 32;
  NND_REQUIRE(blocks_x <= 0x7fffffffLL, "lookup: too many pixels");
  dim3 grid(static_cast<unsigned>(blocks_x), static_cast<unsigned>((n_planes + a.planes_per_block - 1) / a.planes_per_block));
  dim3 block(32, num_levels);
  if (radius <= 5) {
    const size_t smem = static_cast<size_t>(num_levels) * 16 * 32 * sizeof(float);
    if (radius == 4)
      pyramid_lookup_kernel<4, 9><<<grid, block, smem, stream>>>(a);
    else
      pyramid_lookup_kernel<4, 0><<<grid, block, smem, stream>>>(a);
  } else {
    const size_t smem = static_cast<size_t>(num_levels) * 32 * 32 * sizeof(float);
    pyramid_lookup_kernel<8, 0><<<grid, block, smem, stream>>>(a);
  }
  return check_launch("pyramid_lookup_kernel");
}

}  // namespace nnd

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
    NND_REQUIRE(width[l] >= 2, "lookup_indices: level %d has width %d; linear_sampler needs width >= 2", l, width[l]);
    w[l] = width[l];
  }
  const long long n_pix = static_cast<long long>(B) * H * W1;
  const long long total = n_pix * (2 * radius + 1) * num_levels;
  const int threads = 256;
  const long long want = (total + threads - 1) / threads;
  const int blocks = static_cast<int>(want < 148LL * 32 ? want : 148LL * 32);
  nnd::lookup_indices_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      nullptr, coords, n_pix, num_levels, radius, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], idx0, idx1);
  return nnd::check_launch("lookup_indices_kernel");
}

}  // extern "C"
