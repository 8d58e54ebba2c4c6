// inst_packed.cu — one group of instantiations of the generic packed kernel per translation unit.
// Compiled by build.py with  -DSR_TU_T=double|float -DSR_TU_NT=<threads> -DSR_TU_MINB=<n> -DSR_TU_GROUP=<g>:
//   group 0: lean (FP32 only; the FP64 lean path is rod_kernel_lean.cuh)   1: SoftPendulum3D (filter + moving base)
//   group 2: plane contact and the muscle-torque forcings                    3: multi-rod assemblies
//   group 4: tapered rods (FP64; CTA sizes 384 and 1024 only)       5: tapered assembly + COOMM muscle layers (384)
#include <atomic>
#include "launch.cuh"
#include "rod_kernel_packed.cuh"

namespace sr {

template <typename T, int NT, int MINB, bool LAPLACE, bool MOVING, bool CONTACT, bool MULTI, bool TORQUE, bool FASTONLY, bool VARY>
cudaError_t launch_packed_kernel(const RodArgs<T> &A, int rods_per_cta, int grid, cudaStream_t s) {
  const size_t smem = (size_t)packed_smem_words(NT, MULTI, TORQUE) * sizeof(T);
  auto kern = rod_packed_kernel<T, NT, MINB, LAPLACE, MOVING, CONTACT, MULTI, TORQUE, FASTONLY, VARY>;
  // the opt-in above 48 KB is a per-device attribute of the function: one bit per device ordinal
  static std::atomic<unsigned long long> opted{0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ULL << (dev & 63);
  if (!(opted.load(std::memory_order_relaxed) & bit)) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    opted.fetch_or(bit, std::memory_order_relaxed);
  }
  kern<<<grid, NT, smem, s>>>(A, rods_per_cta);
  return cudaGetLastError();
}

// tapered assembly + COOMM muscle layers with per-element activations (group 5, 384 threads)
template <int NT> cudaError_t launch_packed_lmus_kernel(const RodArgs<double> &A, int rods_per_cta, int grid, cudaStream_t s) {
  const size_t smem = (size_t)packed_smem_words(NT, true, false, true) * sizeof(double);
  // (no plane under these assemblies: sr_create rejects contact_on with muscle_layers_on, so the contact code stays out)
  auto kern = rod_packed_kernel<double, NT, 1, false, false, false, true, false, false, true, true>;
  static std::atomic<unsigned long long> opted{0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ULL << (dev & 63);
  if (!(opted.load(std::memory_order_relaxed) & bit)) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    opted.fetch_or(bit, std::memory_order_relaxed);
  }
  kern<<<grid, NT, smem, s>>>(A, rods_per_cta);
  return cudaGetLastError();
}

#define SR_I(L, M, C, MU, TQ, F) \
  template cudaError_t launch_packed_kernel<SR_TU_T, SR_TU_NT, SR_TU_MINB, L, M, C, MU, TQ, F, false>(const RodArgs<SR_TU_T> &, int, int, cudaStream_t);
#define SR_I_VARY(C, MU) \
  template cudaError_t launch_packed_kernel<SR_TU_T, SR_TU_NT, SR_TU_MINB, false, false, C, MU, false, false, true>(const RodArgs<SR_TU_T> &, int, int, cudaStream_t);

// the fast-only / fallback pair exists for FP64 (every group) and for the FP32 lean kernel
#if SR_TU_F64
#define SR_I_FAST(L, M, C, MU, TQ) SR_I(L, M, C, MU, TQ, true)
#else
#define SR_I_FAST(L, M, C, MU, TQ)
#endif

#if SR_TU_GROUP == 0
SR_I(false, false, false, false, false, false)
SR_I(false, false, false, false, false, true)
#elif SR_TU_GROUP == 1
SR_I(true, true, false, false, false, false)
SR_I_FAST(true, true, false, false, false)
#elif SR_TU_GROUP == 2
SR_I(false, false, true, false, false, false)
SR_I_FAST(false, false, true, false, false)
SR_I(false, false, true, false, true, false)
SR_I_FAST(false, false, true, false, true)
#elif SR_TU_GROUP == 3
SR_I(false, false, true, true, false, false)
SR_I_FAST(false, false, true, true, false)
#elif SR_TU_GROUP == 4   // tapered rods (per-element constants from HBM): single rod and assembly, safe variants
SR_I_VARY(true, false)
SR_I_VARY(true, true)
#elif SR_TU_GROUP == 5   // tapered assembly with the muscle layers (longitudinal + transverse, per-element activations)
template cudaError_t launch_packed_lmus_kernel<SR_TU_NT>(const RodArgs<double> &, int, int, cudaStream_t);
#endif

}  // namespace sr
