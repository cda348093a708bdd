// One-CTA-per-problem log-marginal-likelihood kernels: dispatch on the kernel family; the template instances live in
// fit_launch_nu1 / nu3 / nu5 / rbf.cu (one translation unit per family, they compile side by side).
#include "fit_launch.h"

namespace aoed {

struct LmlSmallArgs;
#define AOED_DECL(SUFFIX)                                                                                   \
    cudaError_t launch_lml_blk_##SUFFIX(int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s);         \
    cudaError_t launch_lml_small_##SUFFIX(int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s);
AOED_DECL(nu1) AOED_DECL(nu3) AOED_DECL(nu5) AOED_DECL(rbf)
#undef AOED_DECL

cudaError_t launch_lml_blk(int nu, int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s) {
    switch (nu) {
        case 1: return launch_lml_blk_nu1(DP, a, n_prob, s);
        case 3: return launch_lml_blk_nu3(DP, a, n_prob, s);
        case 5: return launch_lml_blk_nu5(DP, a, n_prob, s);
        default: return launch_lml_blk_rbf(DP, a, n_prob, s);
    }
}

cudaError_t launch_lml_small(int nu, int DP, const LmlSmallArgs& a, int n_prob, cudaStream_t s) {
    switch (nu) {
        case 1: return launch_lml_small_nu1(DP, a, n_prob, s);
        case 3: return launch_lml_small_nu3(DP, a, n_prob, s);
        case 5: return launch_lml_small_nu5(DP, a, n_prob, s);
        default: return launch_lml_small_rbf(DP, a, n_prob, s);
    }
}

}  // namespace aoed
