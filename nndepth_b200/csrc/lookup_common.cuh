// Device helpers shared by the lookup kernels (lookup.cu, lookup_ws.cu): the bit-exact tap arithmetic of
// linear_sampler (raft_stereo/utils.py:4-27 of the reference) and the tcgen05 / mbarrier plumbing.
#pragma once

#include "common.cuh"

namespace nnd {

struct LookupArgs {
  ConstPyramid src[2];  // [0] = feature correlation, [1] = geometry volume (IGEV only)
  const float* coords;
  float* out;
  int hw;           // H * W1 (pixels per image)
  int G;            // planes (groups) per pixel and source
  int n_src;        // 1 or 2
  int num_levels;
  int radius;
  int mode;              // 0: channel = l*(S*G*T) + s*(G*T) + g*T + k ; 1: GroupCorrBlock1D view quirk
  int planes_per_block;  // chunk of the S*G planes handled by one block (blockIdx.y selects it)
  int vec;               // 1: pitches % 4 == 0 and bases 16-byte aligned -> float4 loads
};

// One tap of linear_sampler: position t (fp32, reference op order), neighbours i0 <= i1, lerp weights.
struct Tap {
  int i0, i1;
  float coef, one_minus;
};

struct LevelScale {
  float span;      // w2 - 1
  float inv_span;  // RN(1 / span)
  float inv_pow2;  // 1 / 2**level (exact)
};

// x / span, correctly rounded, for x in [-1, span + 1]:  q = RN(x*y), r = x - q*span (exact, FMA),
// q' = RN(q + r*y) with y = RN(1/span) is the IEEE quotient (Markstein's theorem; span is a small
// positive integer, so y is never the all-ones-significand exception).  Inputs outside [-1, span+1]
// are clamped first, which cannot change clamp(x/span, 0, 1); NaN becomes -1 (-> t = 0).
// The theorem needs the residual r free of underflow, i.e. |x| >= ~2^-100.  For smaller non-zero |x|
// the interpolated VALUE is unaffected (t is then 0 or a denormal: either way the result is row[0]
// exactly), but ceil(t) could differ; EXACT_TINY (the index-reporting kernel) therefore routes those
// inputs through the generic IEEE division.
template <bool EXACT_TINY>
__device__ __forceinline__ float sampler_quotient(float x, const LevelScale& s) {
  x = fminf(fmaxf(x, -1.0f), s.span + 1.0f);
  if (EXACT_TINY && fabsf(x) < 1e-30f) return __fdiv_rn(x, s.span);
  const float q = __fmul_rn(x, s.inv_span);
  const float r = __fmaf_rn(-q, s.span, x);
  return __fmaf_rn(r, s.inv_span, q);
}

__device__ __forceinline__ LevelScale level_scale(int width, int lvl, float centre) {
  LevelScale s;
  s.span = static_cast<float>(width - 1);
  s.inv_span = __frcp_rn(s.span);
  s.inv_pow2 = 1.0f / static_cast<float>(1 << lvl);
  (void)centre;
  return s;
}

template <bool EXACT_TINY = false>
__device__ __forceinline__ Tap make_tap(int k, int r, float centre, const LevelScale& s) {
  // dx + coords / 2**i (cost_volume.py:44-46): dx = k - r is an exact small integer
  const float x = __fadd_rn(static_cast<float>(k - r), centre);
  // clamp(x / (w2-1), 0, 1) * (w2-1)   (utils.py:16-18); __saturatef maps NaN to 0
  const float t = __fmul_rn(__saturatef(sampler_quotient<EXACT_TINY>(x, s)), s.span);
  const float f0 = floorf(t);
  Tap tap;
  tap.i0 = static_cast<int>(f0);
  const bool whole = (t == f0);
  tap.i1 = tap.i0 + (whole ? 0 : 1);                        // ceil(t)
  const float f1 = whole ? f0 : __fadd_rn(f0, 1.0f);        // float(idx1), exact
  tap.coef = __fsub_rn(f1, t);                              // coef = idx1 - t      (utils.py:26)
  tap.one_minus = __fsub_rn(1.0f, tap.coef);                // (1 - coef)           (utils.py:27)
  return tap;
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t y;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(y) : "f"(x));
  return y;
}


namespace umma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // The wait itself is a two-instruction loop (try_wait suspends the warp in hardware for up to the time hint, so a
  // waiting role does not take issue slots from the working ones); every 4096 misses the outer loop checks a watchdog:
  // a protocol bug must fault, not hang the GPU.
  uint32_t done = 0;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 4096;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "@p bra LAB_DONE;\n\t"
        "sub.u32 n, n, 1;\n\t"
        "setp.ne.u32 p, n, 0;\n\t"
        "@p bra LAB_WAIT;\n\t"
        "mov.u32 %0, 0;\n\t"
        "bra LAB_OUT;\n\t"
        "LAB_DONE:\n\t"
        "mov.u32 %0, 1;\n\t"
        "LAB_OUT:\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000u)
        : "memory");
    if (!done) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000LL) __trap();
    }
  } while (!done);
}
// K-major operand without swizzle: 16-byte K chunks of 8 consecutive rows form a 128-byte core matrix;
// lbo = distance between the two K chunks of a k-step, sbo = distance between 8-row groups.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;   // layout type 0: no swizzle
}
// element (row r, column k) of an operand region whose k-steps are blocks of `kstep_bytes`
__device__ __forceinline__ uint32_t operand_offset(int r, int k, int kstep_bytes) {
  return static_cast<uint32_t>((k >> 3) * kstep_bytes + (r >> 3) * 256 + ((k >> 2) & 1) * 128 + (r & 7) * 16 + (k & 3) * 4);
}
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
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
}  // namespace umma

}  // namespace nnd
