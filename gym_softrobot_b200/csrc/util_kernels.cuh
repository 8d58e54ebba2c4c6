// util_kernels.cuh — measurement / self-test kernels (not on the hot path); included by softrod_api.cu only.
#pragma once
#include "rod_math.cuh"

namespace sr {

// register-resident DFMA chains: the FP64 roofline denominator
__global__ void dfma_peak_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same probe with three distinct 64-bit REGISTER operands per DFMA (what the rod kernel issues): shows
// whether operand delivery from the register file, not the FMA units, caps the FP64 issue rate
__global__ void dfma_peak_regs_kernel(double *out, const double *in, int iters) {
  double a[8], b[8], c[8];
  for (int i = 0; i < 8; i++) {
    a[i] = in[(threadIdx.x + i) & 63]; b[i] = in[(threadIdx.x + 8 + i) & 63]; c[i] = in[(threadIdx.x + 16 + i) & 63];
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b[i], c[i]);
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = fma(b[i], c[(i + 1) & 7], a[(i + 3) & 7]);
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i] + b[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// sr_selftest_reciprocals: max relative error of rsqrt_nr / rcp_nr against the IEEE-rounded results
__global__ void reciprocal_selftest_kernel(int n, double lo, double hi, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double e0 = 0.0, e1 = 0.0;
  if (i < n) {
    const double x = lo * exp(log(hi / lo) * (double)i / (double)(n - 1));
    const double r0 = 1.0 / sqrt(x), r1 = 1.0 / x;           // correctly rounded sqrt and divisions
    e0 = fabs(rsqrt_nr(x) / r0 - 1.0);
    e1 = fabs(rcp_nr(x) / r1 - 1.0);
  }
  // errors are non-negative doubles: their bit patterns order like unsigned integers
  atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(e0));
  atomicMax(reinterpret_cast<unsigned long long *>(out + 1), (unsigned long long)__double_as_longlong(e1));
}

// issue-to-use latencies of the instructions the substep's dependency chain is made of: one warp, 1024 dependent
// operations each, cycles per operation.  out: [0] DFMA (register operands)  [1] DADD  [2] MUFU.RSQ64H + DFMA
// [3] shared-memory store -> barrier -> load round trip  [4] DFMA with a constant-bank multiplicand
__global__ void latency_probe_kernel(double *out, double seed) {
  __shared__ double sh[64];
  const int N = 1024;
  double a = seed + threadIdx.x * 1e-12, b = 1.0 + 1e-13, c = 1e-14;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = fma(a, b, c);
  long long t1 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = a + c;
  long long t2 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = fma(y, c, a); }
  long long t3 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) { sh[threadIdx.x] = a; __syncthreads(); a = sh[threadIdx.x ^ 1] + c; __syncthreads(); }
  long long t4 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = fma(a, 1.0000000000001, c);
  long long t5 = clock64();
  if (threadIdx.x == 0) {
    out[0] = (double)(t1 - t0) / N; out[1] = (double)(t2 - t1) / N; out[2] = (double)(t3 - t2) / N;
    out[3] = (double)(t4 - t3) / N; out[4] = (double)(t5 - t4) / N; out[5] = a;
  }
}

}  // namespace sr
