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
// rsqrt() / __drcp_rn().  Checked against the IEEE results by sr_selftest_reciprocals
// (tests/test_parity_gpu.py::test_newton_reciprocals_are_one_ulp).
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y = rsqrt_approx(x);
  double h = x * y;
  double e = fma(-h, y, 1.0);              // 1 - x y^2
  double p = fma(0.375, e, 0.5) * e;       // e/2 + 3 e^2/8
  return fma(y, p, y);
}
__device__ __forceinline__ float rsqrt_nr(float x) {   // rsqrtf is ~2 ulp; one Newton step makes it ~0.5 ulp
  float y = rsqrtf(x);
  float e = fmaf(-x * y, y, 1.0f);
  return fmaf(y, 0.5f * e, y);
}
__device__ __forceinline__ double rcp_nr(double x) {
  double y = rcp_approx(x);
  double e = fma(-x, y, 1.0);              // 1 - x y
  return fma(y, fma(e, e, e), y);          // y (1 + e + e^2)
}
__device__ __forceinline__ float rcp_nr(float x) { return __frcp_rn(x); }

// Polynomial maps below are degree-minimal interpolants at Chebyshev nodes, fitted in
// 60-digit arithmetic (mpmath) and verified to <= 1 ulp (2.2e-16) on the stated range.
// The coefficients travel in the kernel-parameter constant bank (PolyCoef inside RodArgs),
// so every Horner/Estrin step is one DFMA with a c[0][..] operand: no register, no move.

// A = sin(t)/t and B = (1-cos t)/t^2 as functions of q = t^2, valid for q <= 0.25
// (|rotation| <= 0.5 rad per kinematic update); ascending powers.
#define SR_COEF_SINC {1.0, -0.16666666666666116, 0.00833333333307693, -0.00019841269403598734, \
                      2.7556981477681245e-06, -2.4931934029233215e-08}
#define SR_COEF_COSC {0.5, -0.041666666666666276, 0.0013888888888705666, -2.4801586988836374e-05, \
                      2.7557077887523167e-07, -2.0790894217124703e-09}
constexpr double kSmallRotQ = 0.25;

// g(u) = theta / sin(theta) with u = sin^2(theta/2) = (1 - cos theta)/2, valid for
// u <= 0.25, i.e. up to 60 degrees of bending between neighbouring elements (a rod bent
// to a radius of one element length); ascending powers, degree 13.
#define SR_COEF_BEND {1.0, 0.6666666666667322, 0.5333333333163501, \
                      0.4571428588719633, 0.40634911476368524, 0.3694112649313803, \
                      0.3409332932493519, 0.3190722876223924, 0.29180058074110277, \
                      0.33511256175349097, 0.03513915536852488, 0.9776555580022731, \
                      -1.1132861271784307, 1.5537792943093436}
constexpr double kSmallBendU = 0.25;

// exp(z) for |z| <= 0.01: c^(e) = c * exp((e-1) ln c) in the analytical damper; degree 5.
#define SR_COEF_EXP {1.0, 1.0, 0.49999999999218747, 0.1666666666655506, 0.041666875000418525, \
                     0.008333363095284598}
constexpr double kSmallExpZ = 1.0e-2;

// Narrow-domain versions for the fast-only kernel (rod_kernel_packed.cuh, FASTONLY): an env whose arguments
// leave these ranges is re-run by the safe kernel with the wide maps above, so the common case can use
// lower degrees: q <= 0.01 (0.1 rad per kinematic update), u <= 0.04 (23 degrees between neighbouring
// elements), |z| <= 2.5e-4.  Same fitting procedure, <= 1 ulp on the stated range.
#define SR_COEF_SINC3 {0.9999999999999998, -0.16666666666597785, 0.008333332988923206, -0.00019835759066305837}
#define SR_COEF_COSC3 {0.5, -0.04166666666659778, 0.0013888888544469368, -2.4796076411816044e-05}
#define SR_COEF_BEND7 {0.9999999999999999, 0.6666666666668902, 0.5333333332161609, 0.4571428805086313, \
                       0.4063469227387624, 0.3695291784110767, 0.33747354927661083, 0.3708210729102131}
#define SR_COEF_EXP3 {1.0, 1.0, 0.5000000026041667, 0.1666666671875}
// actuated arms on the plane reach 25-30 degrees per element under random actions: their fast-only kernel uses
// u <= 0.1 (37 degrees), degree 9
#define SR_COEF_BEND9 {0.9999999999999999, 0.666666666666836, 0.5333333332775754, 0.4571428642472279, \
                       0.40634874789975095, 0.3694252997188034, 0.34061379942887826, 0.32344955229361544, \
                       0.2573124543735613, 0.46543435203468897}
// theta'/sin(theta') as a function of w2 = |axial(R - R^T)|^2 = 4 sin^2(theta), theta' = acos(cos(theta) - 1e-10) (the
// reference's guarded angle), for w2 <= 0.6144 (the same 23.07 degrees as kNarrowBendU); degree 9, 2.8e-16 fit error
// (scripts/fit_poly.py bendw 0.6144 9).  The lean kernel needs no trace of R with this map.
#define SR_COEF_BENDW {1.0000000000333331, 0.041666666670077346, 0.004687499996336051, 0.0006975447286339237, \
                       0.00011867857307313035, 2.185318201165154e-05, 4.217099326139182e-06, 8.952330387027554e-07, \
                       1.2044932536293865e-07, 7.495667373584703e-08}
constexpr double kNarrowBendW2 = 0.6144;
// the same map on w2 <= 1.449 (37 degrees per element: actuated arms on the plane, where random +-22 actions push a few
// per cent of the arms past 23 degrees every step), degree 13, 5.9e-16 (scripts/fit_poly.py bendw 1.449 13)
#define SR_COEF_BENDW_MID {1.000000000033333, 0.04166666667010045, 0.004687499995989267, 0.0006975447215852305, \
                           0.00011867875407192913, 2.1851702201916664e-05, 4.22290674826398e-06, 8.84413715124705e-07, \
                           1.2412692505191034e-07, 9.612245287938274e-08, -3.7674586576030154e-08, 2.5214678884575354e-08, \
                           -6.966131831983577e-09, 1.2127546473309148e-09}
constexpr double kMidBendW2 = 1.449;
// sin(t)/t = 1 + q g(q) and (1 - cos t)/t^2 = 1/2 + q h(q), q = t^2 <= kNarrowRotQ: degree-2 g, h with the leading
// constants exact (instruction immediates in the lean kernel); 8.6e-16 / 1.7e-16 relative on the range
#define SR_COEF_SINCG {-0.16666666666658056, 0.008333333178343767, -0.0001983713666611297}
#define SR_COEF_COSCH {-0.04166666666665805, 0.0013888888733895929, -2.4797454055979264e-05}
// the lean kernel's c_w^e = c_w exp(z) drops z^3/6: |z| <= 2e-5 keeps that below 1.4e-15
constexpr double kLeanExpZ = 2.0e-5;
constexpr double kLeanExpZc = 1.0e-2;   // degree-6 variant: z^7 / 5040 < 2e-18 (the generic kernel's range)
constexpr double kNarrowRotQ = 0.01, kNarrowBendU = 0.04, kMidBendU = 0.1, kNarrowExpZ = 2.5e-4;

template <typename T> struct PolyCoef {
  T sinc[6], cosc[6], bend[14], expz[6];
  T sinc3[4], cosc3[4], bend7[8], exp3[4], bend9[10];
};

template <typename T> __device__ __forceinline__ void sinc_cosc(const PolyCoef<T> &C, T q, T &A, T &B) {
  T a = fma(C.sinc[5], q, C.sinc[4]);
  T b = fma(C.cosc[5], q, C.cosc[4]);
  a = fma(a, q, C.sinc[3]); b = fma(b, q, C.cosc[3]);
  a = fma(a, q, C.sinc[2]); b = fma(b, q, C.cosc[2]);
  a = fma(a, q, C.sinc[1]); b = fma(b, q, C.cosc[1]);
  A = fma(a, q, C.sinc[0]); B = fma(b, q, C.cosc[0]);
}

// Estrin evaluation of the degree-13 bend polynomial: 7 independent first-level FMAs and a
// depth of 5 instead of a 13-deep Horner chain (the kernel is bound by dependent-issue latency).
template <typename T> __device__ __forceinline__ T theta_over_sin(const PolyCoef<T> &C, T u) {
  const T *c = C.bend;
  T u2 = u * u;
  T p0 = fma(c[1], u, c[0]), p1 = fma(c[3], u, c[2]), p2 = fma(c[5], u, c[4]), p3 = fma(c[7], u, c[6]);
  T p4 = fma(c[9], u, c[8]), p5 = fma(c[11], u, c[10]), p6 = fma(c[13], u, c[12]);
  T u4 = u2 * u2;
  T q0 = fma(p1, u2, p0), q1 = fma(p3, u2, p2), q2 = fma(p5, u2, p4);
  T u8 = u4 * u4;
  T r0 = fma(q1, u4, q0), r1 = fma(p6, u4, q2);
  return fma(r1, u8, r0);
}

template <typename T> __device__ __forceinline__ T exp_small(const PolyCoef<T> &C, T z) {
  T p = fma(C.expz[5], z, C.expz[4]);
  p = fma(p, z, C.expz[3]);
  p = fma(p, z, C.expz[2]);
  p = fma(p, z, C.expz[1]);
  return fma(p, z, C.expz[0]);
}

template <typename T> __device__ __forceinline__ void sinc_cosc_narrow(const PolyCoef<T> &C, T q, T &A, T &B) {
  T a = fma(C.sinc3[3], q, C.sinc3[2]);
  T b = fma(C.cosc3[3], q, C.cosc3[2]);
  a = fma(a, q, C.sinc3[1]); b = fma(b, q, C.cosc3[1]);
  A = fma(a, q, C.sinc3[0]); B = fma(b, q, C.cosc3[0]);
}

template <typename T> __device__ __forceinline__ T theta_over_sin_narrow(const PolyCoef<T> &C, T u) {
  const T *c = C.bend7;
  T u2 = u * u;
  T p0 = fma(c[1], u, c[0]), p1 = fma(c[3], u, c[2]), p2 = fma(c[5], u, c[4]), p3 = fma(c[7], u, c[6]);
  T u4 = u2 * u2;
  T q0 = fma(p1, u2, p0), q1 = fma(p3, u2, p2);
  return fma(q1, u4, q0);
}

template <typename T> __device__ __forceinline__ T theta_over_sin_mid(const PolyCoef<T> &C, T u) {
  const T *c = C.bend9;
  T u2 = u * u;
  T p0 = fma(c[1], u, c[0]), p1 = fma(c[3], u, c[2]), p2 = fma(c[5], u, c[4]), p3 = fma(c[7], u, c[6]), p4 = fma(c[9], u, c[8]);
  T u4 = u2 * u2;
  T q0 = fma(p1, u2, p0), q1 = fma(p3, u2, p2);
  T u8 = u4 * u4;
  return fma(p4, u8, fma(q1, u4, q0));
}

template <typename T> __device__ __forceinline__ T exp_narrow(const PolyCoef<T> &C, T z) {
  T p = fma(C.exp3[3], z, C.exp3[2]);
  p = fma(p, z, C.exp3[1]);
  return fma(p, z, C.exp3[0]);
}

}  // namespace sr
