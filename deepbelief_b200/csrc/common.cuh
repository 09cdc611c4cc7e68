// Shared device helpers for the deepbelief_b200 kernels (sm_100a only).
//
// Nothing here is derived from the reference (which has no native code);
// the RNG restates TensorFlow's documented Philox4x32-10 -> uniform / Box-Muller
// construction [TF, not vendored] that rbm.py:202 / bgrbm.py:108 rely on.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#if !defined(__CUDA_ARCH__) || (__CUDA_ARCH__ >= 1000)
#define DB_SM100 1
#endif

namespace db {

// ----------------------------------------------------------------------------------------------
// Philox4x32-10, counter based. One call yields 4 words.
// Counter convention used by EVERY sampling kernel (and by oracle/philox.py):
//   c0 = unit (column) index, c1 = global_row >> 2, c2 = draw_id low, c3 = draw_id high
//   key = (seed low, seed high);  word j of the result belongs to global_row & 3 == j.
// Keyed by the GLOBAL row so a row-sharded batch draws the same numbers on any GPU count.
// ----------------------------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                      uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
  uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
  uint32_t n0 = hi1 ^ c1 ^ k0;
  uint32_t n1 = lo1;
  uint32_t n2 = hi0 ^ c3 ^ k1;
  uint32_t n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += W0; k1 += W1;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// uniform in [0,1) with 23 random mantissa bits: bits(1.0f) | (x >> 9), minus 1 [TF, not vendored]
__host__ __device__ __forceinline__ float u32_to_uniform(uint32_t x) {
  uint32_t b = 0x3f800000u | (x >> 9);
#ifdef __CUDA_ARCH__
  return __uint_as_float(b) - 1.0f;
#else
  union { uint32_t u; float f; } cv; cv.u = b; return cv.f - 1.0f;
#endif
}

// Box-Muller on two words -> two standard normals [TF, not vendored: BoxMullerFloat]
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  float u1 = fmaxf(u32_to_uniform(a), 1.0e-7f);
  float v1 = 6.28318530717958647692f * u32_to_uniform(b);
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(v1, &s, &c);
  n0 = r * s; n1 = r * c;
}

struct RngArgs {
  uint32_t seed_lo, seed_hi;   // key
  uint32_t draw_lo, draw_hi;   // c2, c3: one draw id per sampling call
  uint32_t row_offset;         // global index of local row 0 (row-sharded batches)
  const unsigned long long* clock;  // optional device clock: draw id += clock[0] (db_rng_t.clock), resolved by rng_resolved
};

// the draw id a kernel launched with a device clock uses (SIMT kernels call this once per thread)
__device__ __forceinline__ RngArgs rng_resolved(RngArgs r) {
  if (r.clock) {
    const unsigned long long d = (((unsigned long long)r.draw_hi << 32) | r.draw_lo) + __ldg(r.clock);
    r.draw_lo = (uint32_t)d; r.draw_hi = (uint32_t)(d >> 32); r.clock = nullptr;
  }
  return r;
}

// uniform for element (global_row, col)
__device__ __forceinline__ float philox_uniform_at(const RngArgs& r, uint32_t grow, uint32_t col) {
  Philox4 o = philox4x32_10(col, grow >> 2, r.draw_lo, r.draw_hi, r.seed_lo, r.seed_hi);
  uint32_t sel = grow & 3u;
  uint32_t w = sel == 0 ? o.x : sel == 1 ? o.y : sel == 2 ? o.z : o.w;
  return u32_to_uniform(w);
}

// standard normal for element (global_row, col): rows (4i,4i+1) share words (x,y), rows (4i+2,4i+3) share (z,w)
__device__ __forceinline__ float philox_normal_at(const RngArgs& r, uint32_t grow, uint32_t col) {
  Philox4 o = philox4x32_10(col, grow >> 2, r.draw_lo, r.draw_hi, r.seed_lo, r.seed_hi);
  float n0, n1;
  if (grow & 2u) box_muller(o.z, o.w, n0, n1); else box_muller(o.x, o.y, n0, n1);
  return (grow & 1u) ? n1 : n0;
}

// ----------------------------------------------------------------------------------------------
// math
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.0f) + log1pf(__expf(-fabsf(x))); }

// relaxed-Bernoulli (Concrete) sample, rbm.py:243-258: eps0 = 10e-7 = 1e-6
__device__ __forceinline__ float concrete_sample(float p, float u, float inv_temperature) {
  const float e0 = 1.0e-6f;
  float pre = logf(p + e0) - logf(1.0f - p + e0) + logf(u + e0) - logf(1.0f - u + e0);
  return sigmoidf_(inv_temperature * pre);
}

// fp16 hi/lo split of a pre-scaled fp32 value (scale is a power of two, chosen by the caller):
// hi = rn16(x), lo = rn16(x - hi)  ->  hi + lo carries >= 22 significant bits of x.
__device__ __forceinline__ void split_f16(float xs, __half& hi, __half& lo) {
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}

// ---- optimizer rule (rbm.py:107-117, 664-670; TF GradientDescent / Momentum / Adam update formulas [TF, not vendored]) ----
// One element: parameter t, gradient g, slots s1 / s2 (Momentum: accumulator; Adam: m, v). Written with explicit
// round-to-nearest intrinsics so that the stand-alone optimizer kernel, the epilogue of the fused dW kernel and the bias
// kernel produce the same bits whatever the compiler would otherwise contract into FMAs around them.
// kind: 0 SGD, 1 Momentum, 2 Adam (DB_OPT_*). lr_t: bias-corrected Adam step size.
struct OptRule { int kind; float lr, lr_t, momentum, beta1, beta2, eps, weight_decay; };
__device__ __forceinline__ float opt_apply(const OptRule& o, float t, float g, float& s1, float& s2) {
  const float gi = __fmaf_rn(o.weight_decay, t, g);
  if (o.kind == 0) return __fmaf_rn(-o.lr, gi, t);
  if (o.kind == 1) {
    s1 = __fmaf_rn(o.momentum, s1, gi);
    return __fmaf_rn(-o.lr, s1, t);
  }
  s1 = __fmaf_rn(o.beta1, s1, __fmul_rn(1.0f - o.beta1, gi));
  s2 = __fmaf_rn(o.beta2, s2, __fmul_rn(__fmul_rn(1.0f - o.beta2, gi), gi));
  return __fsub_rn(t, __fdiv_rn(__fmul_rn(o.lr_t, s1), __fadd_rn(__fsqrt_rn(s2), o.eps)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace db
