// Shared helpers for the sm_100a kernels of the correlation hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "nndepth_b200.h"

namespace nnd {

// ---- thread-local error string behind nnd_last_error_string() -------------------------------
void set_error(const char* fmt, ...);
nnd_status cuda_fail(cudaError_t e, const char* where);
nnd_status check_launch(const char* kernel);
int sm_count();

#define NND_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      nnd::set_error(__VA_ARGS__);              \
      return NND_ERR_INVALID_ARGUMENT;          \
    }                                           \
  } while (0)

// ---- pyramid descriptor passed by value to kernels ------------------------------------------
struct Pyramid {
  float* ptr[NND_MAX_LEVELS];
  int width[NND_MAX_LEVELS];
  int pitch[NND_MAX_LEVELS];
};

struct ConstPyramid {
  const float* ptr[NND_MAX_LEVELS];
  int width[NND_MAX_LEVELS];
  int pitch[NND_MAX_LEVELS];
};

// build.cu: validate a caller pyramid (widths W2 >> l) / pool levels >= 4 from their predecessors
nnd_status fill_pyramid(Pyramid& pyr, int W2, int num_levels, float* const* level, const int* pitch, bool& vec_ok,
                        const char* who);
nnd_status pool_tail(const Pyramid& pyr, int num_levels, long long rows, cudaStream_t stream);

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- the bit-exact sampler contract (raft_stereo/utils.py:15-21 of the reference) ------------
// t = clamp(x / (w-1), 0, 1) * (w-1), both as separately rounded IEEE fp32 operations.
// __saturatef == clamp to [0,1] (NaN -> 0, so a NaN coordinate reads index 0 instead of faulting).
__device__ __forceinline__ float sampler_position(float x, float span) {
  return __fmul_rn(__saturatef(__fdiv_rn(x, span)), span);
}

// coef*v0 + (1-coef)*v1 with every product/sum rounded (no FMA contraction), utils.py:26-27.
__device__ __forceinline__ float sampler_lerp(float t, float i1f, float v0, float v1) {
  const float coef = __fsub_rn(i1f, t);
  return __fadd_rn(__fmul_rn(coef, v0), __fmul_rn(__fsub_rn(1.0f, coef), v1));
}

// (a + b) / 2 as avg_pool1d(.,2) computes it (sum, then divide by the window size).
__device__ __forceinline__ float pool2(float a, float b) { return __fmul_rn(__fadd_rn(a, b), 0.5f); }

// x / d, correctly rounded, in three instructions (Markstein): q = RN(x*r), e = x - q*d (exact, FMA),
// q' = RN(q + e*r) with r = RN(1/d).  Exact whenever the residual neither under- nor overflows and r is
// not the all-ones-significand exception -- true for the scale divisors (sqrt of small integers) and the
// volume magnitudes here; differs from IEEE division by at most 1 ulp otherwise.
__device__ __forceinline__ float div_rn_fast(float x, float d, float r) {
  const float q = __fmul_rn(x, r);
  const float e = __fmaf_rn(-q, d, x);
  return __fmaf_rn(e, r, q);
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- shared epilogue: 4 consecutive columns n0..n0+3 of one volume row -> levels 0..min(L,4)-1 ----
// Lane ^ 1 must hold columns n0 ^ 4 of the same row (all 32 lanes call this together).
__device__ __forceinline__ void store_row_quad(const Pyramid& pyr, int num_levels, long long row, int n0, float4 v,
                                               bool vec_ok, bool row_ok = true) {
  if (!row_ok) n0 = 0x3fffffff;  // past every level's width: the lane only takes part in the shuffle
  const int w0 = pyr.width[0];
  float* p0 = pyr.ptr[0] + row * pyr.pitch[0] + n0;
  if (n0 < w0) {
    if (vec_ok && n0 + 3 < w0) {
      *reinterpret_cast<float4*>(p0) = v;
    } else {
      p0[0] = v.x;
      if (n0 + 1 < w0) p0[1] = v.y;
      if (n0 + 2 < w0) p0[2] = v.z;
      if (n0 + 3 < w0) p0[3] = v.w;
    }
  }
  float l1a = 0.f, l1b = 0.f, l2 = 0.f;
  if (num_levels > 1) {
    l1a = pool2(v.x, v.y);
    l1b = pool2(v.z, v.w);
    const int w1 = pyr.width[1];
    const int j = n0 >> 1;
    float* p1 = pyr.ptr[1] + row * pyr.pitch[1] + j;
    if (j + 1 < w1) {
      if (vec_ok) *reinterpret_cast<float2*>(p1) = make_float2(l1a, l1b);
      else { p1[0] = l1a; p1[1] = l1b; }
    } else if (j < w1) {
      p1[0] = l1a;
    }
  }
  if (num_levels > 2) {
    l2 = pool2(l1a, l1b);
    const int j = n0 >> 2;
    if (j < pyr.width[2]) pyr.ptr[2][row * pyr.pitch[2] + j] = l2;
  }
  if (num_levels > 3) {
    // level 3 pairs my level-2 value with the neighbouring quad's (lane ^ 1 holds columns n0 ^ 4)
    const float other = __shfl_xor_sync(0xffffffffu, l2, 1);
    const int j = n0 >> 3;
    if ((n0 & 4) == 0 && j < pyr.width[3]) pyr.ptr[3][row * pyr.pitch[3] + j] = pool2(l2, other);
  }
}

// ---- IEEE fp16 packing (round to nearest, saturating instead of inf) ------------------------------
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint2 pack_h4(const float4 v) { return make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w)); }
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  float2 r;
  asm("{.reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(r.x), "=f"(r.y) : "r"(v));
  return r;
}
__device__ __forceinline__ float4 unpack_h4(const uint2 v) {
  const float2 a = unpack_h2(v.x), b = unpack_h2(v.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ---- packed fp32 FMA (sm_100: FFMA2, two fused multiply-adds per issued instruction) -------------------
// acc.{x,y} = fma(a.{x,y}, b.{x,y}, acc.{x,y}), each lane of the pair rounded exactly like a scalar fmaf.
__device__ __forceinline__ void ffma2(float2& acc, const float2 a, const float2 b) {
  unsigned long long A, B, C;
  asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(C) : "l"(A), "l"(B));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(C));
}

}  // namespace nnd
