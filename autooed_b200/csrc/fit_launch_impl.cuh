// Template launchers of the one-CTA-per-problem log-marginal-likelihood kernels; instantiated once per kernel family (nu) in
// fit_launch_nu*.cu so that the 64 instances (2 algorithms x 8 padded widths x 4 families) compile in parallel.
#pragma once
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
}  // namespace aoed

// one translation unit per family defines its pair of entry points
#define AOED_FIT_LAUNCH_FAMILY(SUFFIX, NUVAL)                                                                         \
    namespace aoed {                                                                                                  \
    cudaError_t launch_lml_blk_##SUFFIX(int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s) {                  \
        return launch_blk_dp<NUVAL>(DP, a, n_prob, s);                                                                \
    }                                                                                                                 \
    cudaError_t launch_lml_small_##SUFFIX(int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s) {                \
        return launch_small_dp<NUVAL>(DP, a, n_prob, s);                                                              \
    }                                                                                                                 \
    }
