// Launchers of the template-heavy kernel families, each in its own translation unit so that the library builds in
// parallel (the 8 padded widths x 4 kernel families x value / gradient instances dominate the compile time).
#pragma once
#include <cuda_runtime.h>

#include "fit.cuh"
#include "hessian.cuh"
#include "posterior.cuh"
#include "small_eval.cuh"

namespace aoed {

// launch_xcov.cu
int pick_de(int d, int DP);  // effective width of the thread-per-candidate kernels: DP - 4 when the design fits (see xcov_kernel)
void launch_xcov_any(int nu, int DP, int DE, const XcovArgs& a, dim3 grid, bool grad, cudaStream_t s);
void query_occupancy(int nu, int DP, int DE, int& occ_xcov, int& occ_gradsum);
void launch_gradsum(int DP, int DE, const GradsumArgs& a, dim3 grid, cudaStream_t s);
// launch_fitk.cu
void launch_kmat(int nu, int DP, const KmatArgs& a, dim3 grid, cudaStream_t s);
void launch_lmlgrad(int nu, int DP, const LmlGradArgs& a, dim3 grid, cudaStream_t s);
// launch_small_eval.cu
void launch_small_eval_any(int nu, int DP, const SmallEvalArgs& a, dim3 grid, size_t smem, cudaStream_t s);
// launch_hess.cu
void launch_hess(int nu, int DP, const HessArgs& a, dim3 grid, cudaStream_t s);

}  // namespace aoed
