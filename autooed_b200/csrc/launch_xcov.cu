// See launchers.h.
#include "launchers.h"

namespace aoed {

int pick_de(int d, int DP) { return (DP <= 32 && d <= DP - 4) ? DP - 4 : DP; }

namespace {


template <int NU, int DP, int DE>
void launch_xcov(const XcovArgs& a, dim3 grid, bool grad, cudaStream_t s) {
    if (grad)
        xcov_kernel<NU, DP, true, DE><<<grid, XC_THREADS, 0, s>>>(a);
    else
        xcov_kernel<NU, DP, false, DE><<<grid, XC_THREADS, 0, s>>>(a);
}

template <int DP, int DE>
void launch_xcov_nu(int nu, const XcovArgs& a, dim3 grid, bool grad, cudaStream_t s) {
    switch (nu) {
        case 1: launch_xcov<1, DP, DE>(a, grid, grad, s); break;
        case 3: launch_xcov<3, DP, DE>(a, grid, grad, s); break;
        case 5: launch_xcov<5, DP, DE>(a, grid, grad, s); break;
        default: launch_xcov<-1, DP, DE>(a, grid, grad, s); break;
    }
}

}  // namespace

void launch_xcov_any(int nu, int DP, int DE, const XcovArgs& a, dim3 grid, bool grad, cudaStream_t s) {
    switch (DP * 100 + DE) {
        case 804: launch_xcov_nu<8, 4>(nu, a, grid, grad, s); break;
        case 808: launch_xcov_nu<8, 8>(nu, a, grid, grad, s); break;
        case 1612: launch_xcov_nu<16, 12>(nu, a, grid, grad, s); break;
        case 1616: launch_xcov_nu<16, 16>(nu, a, grid, grad, s); break;
        case 2420: launch_xcov_nu<24, 20>(nu, a, grid, grad, s); break;
        case 2424: launch_xcov_nu<24, 24>(nu, a, grid, grad, s); break;
        case 3228: launch_xcov_nu<32, 28>(nu, a, grid, grad, s); break;
        case 3232: launch_xcov_nu<32, 32>(nu, a, grid, grad, s); break;
        case 4848: launch_xcov_nu<48, 48>(nu, a, grid, grad, s); break;
        case 6464: launch_xcov_nu<64, 64>(nu, a, grid, grad, s); break;
        case 9696: launch_xcov_nu<96, 96>(nu, a, grid, grad, s); break;
        default: launch_xcov_nu<128, 128>(nu, a, grid, grad, s); break;
    }
}

namespace {

template <typename K>
int occ_of(K kernel, int threads) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) n = 1;
    return n;
}

template <int DP, int DE>
void occ_pair(int nu, int& occ_xcov, int& occ_gradsum) {
    switch (nu) {
        case 1: occ_xcov = occ_of(xcov_kernel<1, DP, true, DE>, XC_THREADS); break;
        case 3: occ_xcov = occ_of(xcov_kernel<3, DP, true, DE>, XC_THREADS); break;
        case 5: occ_xcov = occ_of(xcov_kernel<5, DP, true, DE>, XC_THREADS); break;
        default: occ_xcov = occ_of(xcov_kernel<-1, DP, true, DE>, XC_THREADS); break;
    }
    occ_gradsum = occ_of(gradsum_kernel<DP, DE>, XC_THREADS);
}

}  // namespace

void query_occupancy(int nu, int DP, int DE, int& occ_xcov, int& occ_gradsum) {
    switch (DP * 100 + DE) {
        case 804: occ_pair<8, 4>(nu, occ_xcov, occ_gradsum); break;
        case 808: occ_pair<8, 8>(nu, occ_xcov, occ_gradsum); break;
        case 1612: occ_pair<16, 12>(nu, occ_xcov, occ_gradsum); break;
        case 1616: occ_pair<16, 16>(nu, occ_xcov, occ_gradsum); break;
        case 2420: occ_pair<24, 20>(nu, occ_xcov, occ_gradsum); break;
        case 2424: occ_pair<24, 24>(nu, occ_xcov, occ_gradsum); break;
        case 3228: occ_pair<32, 28>(nu, occ_xcov, occ_gradsum); break;
        case 3232: occ_pair<32, 32>(nu, occ_xcov, occ_gradsum); break;
        case 4848: occ_pair<48, 48>(nu, occ_xcov, occ_gradsum); break;
        case 6464: occ_pair<64, 64>(nu, occ_xcov, occ_gradsum); break;
        case 9696: occ_pair<96, 96>(nu, occ_xcov, occ_gradsum); break;
        default: occ_pair<128, 128>(nu, occ_xcov, occ_gradsum); break;
    }
}

namespace {

}  // namespace

void launch_gradsum(int DP, int DE, const GradsumArgs& a, dim3 grid, cudaStream_t s) {
    switch (DP * 100 + DE) {
        case 804: gradsum_kernel<8, 4><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 808: gradsum_kernel<8, 8><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 1612: gradsum_kernel<16, 12><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 1616: gradsum_kernel<16, 16><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 2420: gradsum_kernel<24, 20><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 2424: gradsum_kernel<24, 24><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 3228: gradsum_kernel<32, 28><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 3232: gradsum_kernel<32, 32><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 4848: gradsum_kernel<48, 48><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 6464: gradsum_kernel<64, 64><<<grid, XC_THREADS, 0, s>>>(a); break;
        case 9696: gradsum_kernel<96, 96><<<grid, XC_THREADS, 0, s>>>(a); break;
        default: gradsum_kernel<128, 128><<<grid, XC_THREADS, 0, s>>>(a); break;
    }
}

namespace {


}  // namespace
}  // namespace aoed
