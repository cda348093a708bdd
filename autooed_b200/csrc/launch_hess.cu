// See launchers.h.
#include "launchers.h"

namespace aoed {
namespace {

template <int NU>
void launch_hess_dp(int DP, const HessArgs& a, dim3 grid, cudaStream_t s) {
    switch (DP) {
        case 8: hess_kernel<NU, 8><<<grid, HS_THREADS, 0, s>>>(a); break;
        case 16: hess_kernel<NU, 16><<<grid, HS_THREADS, 0, s>>>(a); break;
        case 24: hess_kernel<NU, 24><<<grid, HS_THREADS, 0, s>>>(a); break;
        case 32: hess_kernel<NU, 32><<<grid, HS_THREADS, 0, s>>>(a); break;
        case 48: hess_kernel<NU, 48><<<grid, HS_THREADS, 0, s>>>(a); break;
        case 64: hess_kernel<NU, 64><<<grid, HS_THREADS, 0, s>>>(a); break;
        case 96: hess_kernel<NU, 96><<<grid, HS_THREADS, 0, s>>>(a); break;
        default: hess_kernel<NU, 128><<<grid, HS_THREADS, 0, s>>>(a); break;
    }
}

}  // namespace

void launch_hess(int nu, int DP, const HessArgs& a, dim3 grid, cudaStream_t s) {
    switch (nu) {
        case 1: launch_hess_dp<1>(DP, a, grid, s); break;
        case 3: launch_hess_dp<3>(DP, a, grid, s); break;
        case 5: launch_hess_dp<5>(DP, a, grid, s); break;
        default: launch_hess_dp<-1>(DP, a, grid, s); break;
    }
}

namespace {


}  // namespace
}  // namespace aoed
