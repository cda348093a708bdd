// One-CTA-per-problem log-marginal-likelihood kernels: their 48 template instances (4 kernels x 6 padded widths x 2
// algorithms) are the slowest part of the build, so they live in their own translation unit and compile next to gp.cu.
#include "fit_launch.h"

#include "fit_small.cuh"
#include "fit_small2.cuh"

namespace aoed {
namespace {

#define FL_CU(call)                        \
    do {                                   \
        cudaError_t e_ = (call);           \
        if (e_ != cudaSuccess) return e_;  \
    } while (0)

template <int NU, int D>
cudaError_t launch_small_one(const LmlSmallArgs& a, int n_prob, size_t sm, cudaStream_t s) {
    FL_CU(cudaFuncSetAttribute(lml_small_kernel<NU, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
    lml_small_kernel<NU, D><<<n_prob, FS_THREADS, sm, s>>>(a);
    return cudaGetLastError();
}

template <int NU, int D>
cudaError_t launch_blk_one(const LmlSmallArgs& a, int n_prob, size_t sm, cudaStream_t s) {
    FL_CU(cudaFuncSetAttribute(lml_blk_kernel<NU, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
    lml_blk_kernel<NU, D><<<n_prob, FS_THREADS, sm, s>>>(a);
    return cudaGetLastError();
}

template <int NU>
cudaError_t launch_small_dp(int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s) {
    const size_t sm = lml_small_smem(a.N, DP);
    switch (DP) {
        case 8: return launch_small_one<NU, 8>(a, n_prob, sm, s);
        case 16: return launch_small_one<NU, 16>(a, n_prob, sm, s);
        case 24: return launch_small_one<NU, 24>(a, n_prob, sm, s);
        case 32: return launch_small_one<NU, 32>(a, n_prob, sm, s);
        case 48: return launch_small_one<NU, 48>(a, n_prob, sm, s);
        case 64: return launch_small_one<NU, 64>(a, n_prob, sm, s);
        case 96: return launch_small_one<NU, 96>(a, n_prob, sm, s);
        default: return launch_small_one<NU, 128>(a, n_prob, sm, s);
    }
}

template <int NU>
cudaError_t launch_blk_dp(int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s) {
    const size_t sm = lml_blk_smem(a.N, DP);
    switch (DP) {
        case 8: return launch_blk_one<NU, 8>(a, n_prob, sm, s);
        case 16: return launch_blk_one<NU, 16>(a, n_prob, sm, s);
        case 24: return launch_blk_one<NU, 24>(a, n_prob, sm, s);
        case 32: return launch_blk_one<NU, 32>(a, n_prob, sm, s);
        case 48: return launch_blk_one<NU, 48>(a, n_prob, sm, s);
        case 64: return launch_blk_one<NU, 64>(a, n_prob, sm, s);
        case 96: return launch_blk_one<NU, 96>(a, n_prob, sm, s);
        default: return launch_blk_one<NU, 128>(a, n_prob, sm, s);
    }
}

}  // namespace

cudaError_t launch_lml_blk(int nu, int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s) {
    switch (nu) {
        case 1: return launch_blk_dp<1>(DP, a, n_prob, s);
        case 3: return launch_blk_dp<3>(DP, a, n_prob, s);
        case 5: return launch_blk_dp<5>(DP, a, n_prob, s);
        default: return launch_blk_dp<-1>(DP, a, n_prob, s);
    }
}

cudaError_t launch_lml_small(int nu, int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s) {
    switch (nu) {
        case 1: return launch_small_dp<1>(DP, a, n_prob, s);
        case 3: return launch_small_dp<3>(DP, a, n_prob, s);
        case 5: return launch_small_dp<5>(DP, a, n_prob, s);
        default: return launch_small_dp<-1>(DP, a, n_prob, s);
    }
}

}  // namespace aoed
