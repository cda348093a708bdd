// Launchers of the one-CTA-per-problem log-marginal-likelihood kernels (defined in fit_launch.cu, its own translation unit).
#pragma once
#include <cuda_runtime.h>

namespace aoed {

struct LmlSmallArgs;

// blocked DMMA kernel (fit_small2.cuh) / rank-1 kernel (fit_small.cuh) for n_prob problems, dispatched on the kernel
// family nu in {1, 3, 5, -1} and the padded design width DP
cudaError_t launch_lml_blk(int nu, int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s);
cudaError_t launch_lml_small(int nu, int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s);

}  // namespace aoed
