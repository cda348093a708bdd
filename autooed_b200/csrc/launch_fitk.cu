// See launchers.h.
#include "launchers.h"

namespace aoed {
namespace {

template <int NU>
void launch_kmat_dp(int DP, const KmatArgs& a, dim3 grid, cudaStream_t s) {
    switch (DP) {
        case 8: kmat_kernel<NU, 8><<<grid, 256, 0, s>>>(a); break;
        case 16: kmat_kernel<NU, 16><<<grid, 256, 0, s>>>(a); break;
        case 24: kmat_kernel<NU, 24><<<grid, 256, 0, s>>>(a); break;
        case 32: kmat_kernel<NU, 32><<<grid, 256, 0, s>>>(a); break;
        case 48: kmat_kernel<NU, 48><<<grid, 256, 0, s>>>(a); break;
        case 64: kmat_kernel<NU, 64><<<grid, 256, 0, s>>>(a); break;
        case 96: kmat_kernel<NU, 96><<<grid, 256, 0, s>>>(a); break;
        default: kmat_kernel<NU, 128><<<grid, 256, 0, s>>>(a); break;
    }
}

}  // namespace

void launch_kmat(int nu, int DP, const KmatArgs& a, dim3 grid, cudaStream_t s) {
    switch (nu) {
        case 1: launch_kmat_dp<1>(DP, a, grid, s); break;
        case 3: launch_kmat_dp<3>(DP, a, grid, s); break;
        case 5: launch_kmat_dp<5>(DP, a, grid, s); break;
        default: launch_kmat_dp<-1>(DP, a, grid, s); break;
    }
}

namespace {

template <int NU>
void launch_lmlgrad_dp(int DP, const LmlGradArgs& a, dim3 grid, cudaStream_t s) {
    switch (DP) {
        case 8: lml_grad_kernel<NU, 8><<<grid, 256, 0, s>>>(a); break;
        case 16: lml_grad_kernel<NU, 16><<<grid, 256, 0, s>>>(a); break;
        case 24: lml_grad_kernel<NU, 24><<<grid, 256, 0, s>>>(a); break;
        case 32: lml_grad_kernel<NU, 32><<<grid, 256, 0, s>>>(a); break;
        case 48: lml_grad_kernel<NU, 48><<<grid, 256, 0, s>>>(a); break;
        case 64: lml_grad_kernel<NU, 64><<<grid, 256, 0, s>>>(a); break;
        case 96: lml_grad_kernel<NU, 96><<<grid, 256, 0, s>>>(a); break;
        default: lml_grad_kernel<NU, 128><<<grid, 256, 0, s>>>(a); break;
    }
}

}  // namespace

void launch_lmlgrad(int nu, int DP, const LmlGradArgs& a, dim3 grid, cudaStream_t s) {
    switch (nu) {
        case 1: launch_lmlgrad_dp<1>(DP, a, grid, s); break;
        case 3: launch_lmlgrad_dp<3>(DP, a, grid, s); break;
        case 5: launch_lmlgrad_dp<5>(DP, a, grid, s); break;
        default: launch_lmlgrad_dp<-1>(DP, a, grid, s); break;
    }
}

namespace {


}  // namespace
}  // namespace aoed
