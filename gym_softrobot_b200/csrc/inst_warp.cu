// inst_warp.cu — warp-per-rod kernel with faithful math (rod_kernels.cuh): the parity build (SR_MATH_FAITHFUL).
#include "launch.cuh"

namespace sr {

template <typename T, int EPL> cudaError_t launch_warp_faithful(const RodArgs<T> &A, int grid, cudaStream_t s) {
  rod_substeps_kernel<T, EPL, MATH_FAITHFUL, 2><<<grid, WARPS_PER_CTA * 32, 0, s>>>(A);
  return cudaGetLastError();
}

template cudaError_t launch_warp_faithful<double, 1>(const RodArgs<double> &, int, cudaStream_t);
template cudaError_t launch_warp_faithful<double, 2>(const RodArgs<double> &, int, cudaStream_t);
template cudaError_t launch_warp_faithful<double, 4>(const RodArgs<double> &, int, cudaStream_t);

}  // namespace sr
