// See launchers.h.
#include <algorithm>

#include "launchers.h"

namespace aoed {
namespace {

template <int NU, int DP>
void launch_small_eval(const SmallEvalArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
    static thread_local int opted_in_device = -1;  // shared-memory opt-in (up to 105 KB at n_train 256, d 32): once per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (opted_in_device != dev) {
        cudaFuncSetAttribute(small_eval_kernel<NU, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(std::min<size_t>(small_eval_smem(SE_MAX_NPAD, DP), 227 * 1024)));
        opted_in_device = dev;
    }
    small_eval_kernel<NU, DP><<<grid, SE_THREADS, smem, s>>>(a);
}

template <int NU>
void launch_small_eval_dp(int DP, const SmallEvalArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
    switch (DP) {
        case 8: launch_small_eval<NU, 8>(a, grid, smem, s); break;
        case 16: launch_small_eval<NU, 16>(a, grid, smem, s); break;
        case 24: launch_small_eval<NU, 24>(a, grid, smem, s); break;
        case 32: launch_small_eval<NU, 32>(a, grid, smem, s); break;
        case 48: launch_small_eval<NU, 48>(a, grid, smem, s); break;
        case 64: launch_small_eval<NU, 64>(a, grid, smem, s); break;
        case 96: launch_small_eval<NU, 96>(a, grid, smem, s); break;
        default: launch_small_eval<NU, 128>(a, grid, smem, s); break;
    }
}

}  // namespace

void launch_small_eval_any(int nu, int DP, const SmallEvalArgs& a, dim3 grid, size_t smem, cudaStream_t s) {
    switch (nu) {
        case 1: launch_small_eval_dp<1>(DP, a, grid, smem, s); break;
        case 3: launch_small_eval_dp<3>(DP, a, grid, smem, s); break;
        case 5: launch_small_eval_dp<5>(DP, a, grid, smem, s); break;
        default: launch_small_eval_dp<-1>(DP, a, grid, smem, s); break;
    }
}

namespace {

}  // namespace
}  // namespace aoed
