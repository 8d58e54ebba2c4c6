// rod_math.cuh — small-vector helpers and the strength-reduced maps used by the
// fused Cosserat-rod substep kernel (sm_100a).  No tensor cores: nothing on this
// path is a dense contraction; the binding unit is the FP64 FMA pipe.
//
// Algorithm being computed: PyElastica's PositionVerlet / CosseratRod step as
// driven by /root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:183-184
// (SURVEY.md Appendix A).  The "faithful" variants keep the reference's
// operation order and libm calls; the "fast" variants are algebraically equal
// maps evaluated without sqrt/div/sin/cos/acos/pow where the argument is small,
// validated against the CPU oracle to 1e-9 (tests/test_parity_gpu.py).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace sr {

constexpr unsigned FULL = 0xffffffffu;

template <typename T> struct V3 { T x, y, z; };

template <typename T> __device__ __forceinline__ T dot3(const T a[3], const T b[3]) {
  return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0]));
}
template <typename T> __device__ __forceinline__ void cross3(const T a[3], const T b[3], T o[3]) {
  o[0] = fma(a[1], b[2], -(a[2] * b[1]));
  o[1] = fma(a[2], b[0], -(a[0] * b[2]));
  o[2] = fma(a[0], b[1], -(a[1] * b[0]));
}

// ---- math wrappers so one template serves double and float -----------------
__device__ __forceinline__ double rsqrt_(double x) { return rsqrt(x); }
__device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rcp_(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float rcp_(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ void sincos_(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __forceinline__ void sincos_(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ double acos_(double x) { return acos(x); }
__device__ __forceinline__ float acos_(float x) { return acosf(x); }
__device__ __forceinline__ double sin_(double x) { return sin(x); }
__device__ __forceinline__ float sin_(float x) { return sinf(x); }
__device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float pow_(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float exp_(float x) { return expf(x); }
__device__ __forceinline__ double atan_(double x) { return atan(x); }
__device__ __forceinline__ double fabs_(double x) { return fabs(x); }
__device__ __forceinline__ float fabs_(float x) { return fabsf(x); }

// ~2^-22-accurate reciprocal square root: one MUFU.RSQ64H, no Newton steps.  Only used
// for the reference's 1e-14 guard terms, where the correction itself is <= 1e-9.
__device__ __forceinline__ double rsqrt_approx(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) { return rsqrtf(x); }

// sin(t)/t and (1-cos(t))/t^2 as polynomials in q = t^2, valid for q <= 1/16
// (|t| <= 0.25 rad per half step; truncation < 1e-17).  Larger rotations take
// the libm path (warp-uniform branch in the kernel).
template <typename T> __device__ __forceinline__ void sinc_cosc(T q, T &A, T &B) {
  // A = sum (-1)^k q^k/(2k+1)!   B = sum (-1)^k q^k/(2k+2)!
  T a = T(1.0 / 6227020800.0);          // 1/13!
  a = fma(a, q, T(-1.0 / 39916800.0));  // 1/11!
  a = fma(a, q, T(1.0 / 362880.0));     // 1/9!
  a = fma(a, q, T(-1.0 / 5040.0));
  a = fma(a, q, T(1.0 / 120.0));
  a = fma(a, q, T(-1.0 / 6.0));
  A = fma(a, q, T(1.0));
  T b = T(1.0 / 87178291200.0);         // 1/14!
  b = fma(b, q, T(-1.0 / 479001600.0)); // 1/12!
  b = fma(b, q, T(1.0 / 3628800.0));    // 1/10!
  b = fma(b, q, T(-1.0 / 40320.0));
  b = fma(b, q, T(1.0 / 720.0));
  b = fma(b, q, T(-1.0 / 24.0));
  B = fma(b, q, T(0.5));
}
constexpr double kSmallRotQ = 0.0625;

// g(u) = theta / sin(theta) with u = sin^2(theta/2) = (1 - cos theta)/2:
//   g = asin(s) / (s sqrt(1-s^2)), s^2 = u  =  sum_k 4^k (k!)^2/(2k+1)! u^k.
// 12 terms: truncation < 1e-17 for u <= 0.0225 (theta <= 0.30 rad between
// neighbouring elements); beyond that the kernel takes the acos path.
template <typename T> __device__ __forceinline__ T theta_over_sin(T u) {
  // c_{k+1} = c_k * (2k+2)/(2k+3)
  constexpr double c0 = 1.0, c1 = c0 * 2 / 3, c2 = c1 * 4 / 5, c3 = c2 * 6 / 7, c4 = c3 * 8 / 9,
                   c5 = c4 * 10 / 11, c6 = c5 * 12 / 13, c7 = c6 * 14 / 15, c8 = c7 * 16 / 17,
                   c9 = c8 * 18 / 19, c10 = c9 * 20 / 21, c11 = c10 * 22 / 23;
  T g = T(c11);
  g = fma(g, u, T(c10));
  g = fma(g, u, T(c9));
  g = fma(g, u, T(c8));
  g = fma(g, u, T(c7));
  g = fma(g, u, T(c6));
  g = fma(g, u, T(c5));
  g = fma(g, u, T(c4));
  g = fma(g, u, T(c3));
  g = fma(g, u, T(c2));
  g = fma(g, u, T(c1));
  g = fma(g, u, T(c0));
  return g;
}
constexpr double kSmallBendU = 0.0225;

// exp(z) for |z| <= 1e-3 (degree 4, truncation 8e-18): used for c^(e) = c * exp((e-1) ln c)
template <typename T> __device__ __forceinline__ T exp_small(T z) {
  T p = T(1.0 / 24.0);
  p = fma(p, z, T(1.0 / 6.0));
  p = fma(p, z, T(0.5));
  p = fma(p, z, T(1.0));
  return fma(p, z, T(1.0));
}
constexpr double kSmallExpZ = 1.0e-3;

}  // namespace sr
