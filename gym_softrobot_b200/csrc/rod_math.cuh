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

// ~2^-20-accurate seeds: one MUFU each (only the upper 32 bits of the operand are read).
__device__ __forceinline__ double rsqrt_approx(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp_approx(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) { return rsqrtf(x); }
__device__ __forceinline__ float rcp_approx(float x) { return __frcp_rn(x); }

// Full-precision 1/sqrt(x) and 1/x for normal positive x (lengths, dilatations):
// MUFU seed + one third-order Newton step (seed error e ~ 2^-20 -> e^3 ~ 2^-60).
// 5 / 3 FP64 instructions instead of the ~10 / ~8 (plus special-case branch) of
// rsqrt() / __drcp_rn().  Checked against the IEEE results in tests/test_kernels_gpu.py.
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y = rsqrt_approx(x);
  double h = x * y;
  double e = fma(-h, y, 1.0);              // 1 - x y^2
  double p = fma(0.375, e, 0.5) * e;       // e/2 + 3 e^2/8
  return fma(y, p, y);
}
__device__ __forceinline__ float rsqrt_nr(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rcp_nr(double x) {
  double y = rcp_approx(x);
  double e = fma(-x, y, 1.0);              // 1 - x y
  return fma(y, fma(e, e, e), y);          // y (1 + e + e^2)
}
__device__ __forceinline__ float rcp_nr(float x) { return __frcp_rn(x); }

// Polynomial maps below are degree-minimal interpolants at Chebyshev nodes, fitted in
// 60-digit arithmetic (mpmath) and verified to <= 1 ulp (2.2e-16) on the stated range.

// A = sin(t)/t and B = (1-cos t)/t^2 as functions of q = t^2, valid for q <= 0.25
// (|rotation| <= 0.5 rad per kinematic update).  Larger rotations take the libm path.
template <typename T> __device__ __forceinline__ void sinc_cosc(T q, T &A, T &B) {
  T a = T(-2.4931934029233215e-08);
  a = fma(a, q, T(2.7556981477681245e-06));
  a = fma(a, q, T(-0.00019841269403598734));
  a = fma(a, q, T(0.00833333333307693));
  a = fma(a, q, T(-0.16666666666666116));
  A = fma(a, q, T(1.0));
  T b = T(-2.0790894217124703e-09);
  b = fma(b, q, T(2.7557077887523167e-07));
  b = fma(b, q, T(-2.4801586988836374e-05));
  b = fma(b, q, T(0.0013888888888705666));
  b = fma(b, q, T(-0.041666666666666276));
  B = fma(b, q, T(0.5));
}
constexpr double kSmallRotQ = 0.25;

// g(u) = theta / sin(theta) with u = sin^2(theta/2) = (1 - cos theta)/2, valid for
// u <= 0.25, i.e. up to 60 degrees of bending between neighbouring elements (a rod bent
// to a radius of one element length).  Beyond that the kernel takes the acos path.
template <typename T> __device__ __forceinline__ T theta_over_sin(T u) {
  T g = T(0.7004333670912273);
  g = fma(g, u, T(-0.010109211165683595));
  g = fma(g, u, T(0.33627006293385214));
  g = fma(g, u, T(0.2554785051838476));
  g = fma(g, u, T(0.2856689010559272));
  g = fma(g, u, T(0.2993695424665021));
  g = fma(g, u, T(0.3182700383289065));
  g = fma(g, u, T(0.3409918863270117));
  g = fma(g, u, T(0.36940838269940973));
  g = fma(g, u, T(0.4063492060981829));
  g = fma(g, u, T(0.4571428571456908));
  g = fma(g, u, T(0.5333333333333167));
  g = fma(g, u, T(0.6666666666666667));
  return fma(g, u, T(1.0));
}
constexpr double kSmallBendU = 0.25;

// exp(z) for |z| <= 0.01: used for c^(e) = c * exp((e-1) ln c) in the analytical damper
template <typename T> __device__ __forceinline__ T exp_small(T z) {
  T p = T(0.008333363095284598);
  p = fma(p, z, T(0.041666875000418525));
  p = fma(p, z, T(0.1666666666655506));
  p = fma(p, z, T(0.49999999999218747));
  p = fma(p, z, T(1.0));
  return fma(p, z, T(1.0));
}
constexpr double kSmallExpZ = 1.0e-2;

}  // namespace sr
