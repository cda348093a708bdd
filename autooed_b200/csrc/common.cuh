// Shared device helpers for the sm_100a kernels of the GP-surrogate hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aoed {

constexpr double kJitter = 1e-10;          // sklearn GaussianProcessRegressor(alpha=1e-10), gp.py:134
constexpr double kSqrt3 = 1.7320508075688772;
constexpr double kSqrt5 = 2.23606797749979;
constexpr double kInvSqrt2Pi = 0.3989422804014327;
constexpr double kInvSqrt2 = 0.7071067811865476;

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// Fragment ownership (lane = 4*g + q): A[g][q], B[q][g], C[g][2q], C[g][2q+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 16-byte asynchronous global->shared copy (LDGSTS), zero-filled when !pred.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// Kernel profile phi(r) and radial derivative factor (dphi/dr)/r for nu in {1,3,5,-1}
// (gp.py:36-43 and gp.py:181-191; RBF = sklearn kernels.RBF).  r2 = r*r is passed as well so that the
// RBF branch needs no sqrt.  For nu=1 the factor is singular at r=0: safe_divide -> 0 (gp.py:176-179).
template <int NU>
__device__ __forceinline__ void kernel_profile(double r, double r2, double& phi, double& g_over_r) {
    if (NU == 1) {
        double e = exp(-r);
        phi = e;
        g_over_r = (r != 0.0) ? -e / r : 0.0;
    } else if (NU == 3) {
        double s = kSqrt3 * r;
        double e = exp(-s);
        phi = (1.0 + s) * e;
        g_over_r = -3.0 * e;
    } else if (NU == 5) {
        double s = kSqrt5 * r;
        double e = exp(-s);
        phi = (1.0 + s + s * s * (1.0 / 3.0)) * e;  // reciprocal constant: a true divide costs ~10 FP64 instructions per pair
        g_over_r = -5.0 / 3.0 * e * (1.0 + s);
    } else {
        double e = exp(-0.5 * r2);
        phi = e;
        g_over_r = -e;
    }
}

__device__ __forceinline__ double norm_pdf(double z) { return kInvSqrt2Pi * exp(-0.5 * z * z); }
__device__ __forceinline__ double norm_cdf(double z) { return 0.5 * erfc(-z * kInvSqrt2); }
__device__ __forceinline__ double safe_div(double a, double b) { return b != 0.0 ? a / b : 0.0; }  // utils/operand.py:8-12

// Selects `dev` for the scope of one C-ABI call and restores the caller's current device afterwards (the host process
// may be a torch program whose current device must not change under its feet).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace aoed
