// Thompson-sampling acquisition on the GPU (sm_100a, FP64): value / gradient / Hessian of a random-Fourier-feature
// posterior sample  f_o(x) = factor_o * sum_j theta_oj cos(W_oj . x + b_oj)
// (replaces `ThompsonSampling._evaluate`, autooed/mobo/acquisition/ts.py:65-87, and the un-normalisation it applies:
// F -> F * y_scale + y_mean, dF -> dF * y_scale, hF left in normalised-Y units).
//
//   rff_kernel<DP, GRAD>   thread = candidate, spectral points (W, b, theta) of one objective broadcast from shared memory;
//                          M * (d + 2) fused multiply-adds and one sincos per spectral point
//   rff_hess_kernel        one CTA per (candidate, objective), thread per (i, k): -factor * sum_j theta_j cos(t_j) W_ji W_jk
#include <cuda_runtime.h>

#include <string>

#include "../../include/aoed_b200.h"
#include "common.cuh"
#include "rff.cuh"

extern "C" int aoed_detail_fail(int code, const char* msg);

namespace {

using namespace aoed;

int rfail(int code, const std::string& msg) { return aoed_detail_fail(code, msg.c_str()); }
#define RCU(call)                                                                                                    \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess)                                                                                       \
            return rfail(AOED_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + \
                                            std::to_string(__LINE__));                                               \
    } while (0)

struct DBuf {
    void* p = nullptr;
    ~DBuf() {
        if (p) cudaFree(p);
    }
};

template <int DP>
void launch_rff(const RffArgs& a, bool grad, dim3 grid, size_t sm, cudaStream_t s) {
    if (grad)
        rff_kernel<DP, true><<<grid, 128, sm, s>>>(a);
    else
        rff_kernel<DP, false><<<grid, 128, sm, s>>>(a);
}

}  // namespace

extern "C" int aoed_rff_eval(int device, int64_t n_rows, int n_var, int n_obj, int n_feat, const double* h_W, const double* h_b,
                             const double* h_theta, const double* h_factor, const double* h_y_mean, const double* h_y_scale,
                             const double* h_Xn, uint32_t flags, double* h_F, double* h_dF, double* h_hF) {
    if (n_rows < 0 || n_var < 1 || n_obj < 1 || n_feat < 1 || !h_W || !h_b || !h_theta || !h_factor || !h_y_mean || !h_y_scale)
        return rfail(AOED_ERR_INVALID, "bad argument");
    if (n_rows == 0) return AOED_OK;
    if (!h_Xn) return rfail(AOED_ERR_INVALID, "null input");
    if (n_var > 128) return rfail(AOED_ERR_UNSUPPORTED, "n_var > 128 is not supported");
    const int DP = aoed::rff_dp(n_var);
    const size_t sm = (size_t(n_feat) * DP + 2 * size_t(n_feat)) * sizeof(double);
    if (sm > 200 * 1024) return rfail(AOED_ERR_UNSUPPORTED, "too many spectral points for shared memory");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return rfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    RCU(guard_.err);
    const bool grad = (flags & AOED_EVAL_GRAD) != 0 && h_dF, hess = (flags & AOED_EVAL_HESS) != 0 && h_hF;
    const size_t d = n_var, m = n_obj, M = n_feat, C = size_t(n_rows);
    DBuf dX, dP, dF, ddF, dhF;
    const size_t np = m * M * d + 2 * m * M + 3 * m;
    RCU(cudaMalloc(&dX.p, C * d * sizeof(double)));
    RCU(cudaMalloc(&dP.p, np * sizeof(double)));
    double* P = static_cast<double*>(dP.p);
    RffArgs a;
    a.X = static_cast<double*>(dX.p);
    a.W = P; a.b = P + m * M * d; a.theta = a.b + m * M; a.factor = a.theta + m * M; a.y_mean = a.factor + m; a.y_scale = a.y_mean + m;
    RCU(cudaMemcpy(dX.p, h_Xn, C * d * sizeof(double), cudaMemcpyHostToDevice));
    RCU(cudaMemcpy(const_cast<double*>(a.W), h_W, m * M * d * sizeof(double), cudaMemcpyHostToDevice));
    RCU(cudaMemcpy(const_cast<double*>(a.b), h_b, m * M * sizeof(double), cudaMemcpyHostToDevice));
    RCU(cudaMemcpy(const_cast<double*>(a.theta), h_theta, m * M * sizeof(double), cudaMemcpyHostToDevice));
    RCU(cudaMemcpy(const_cast<double*>(a.factor), h_factor, m * sizeof(double), cudaMemcpyHostToDevice));
    RCU(cudaMemcpy(const_cast<double*>(a.y_mean), h_y_mean, m * sizeof(double), cudaMemcpyHostToDevice));
    RCU(cudaMemcpy(const_cast<double*>(a.y_scale), h_y_scale, m * sizeof(double), cudaMemcpyHostToDevice));
    a.F = a.dF = a.hF = nullptr;
    if (h_F) {
        RCU(cudaMalloc(&dF.p, C * m * sizeof(double)));
        a.F = static_cast<double*>(dF.p);
    }
    if (grad) {
        RCU(cudaMalloc(&ddF.p, C * m * d * sizeof(double)));
        a.dF = static_cast<double*>(ddF.p);
    }
    if (hess) {
        RCU(cudaMalloc(&dhF.p, C * m * d * d * sizeof(double)));
        a.hF = static_cast<double*>(dhF.p);
    }
    a.rows = n_rows; a.d = n_var; a.m = n_obj; a.M = n_feat;
    if (h_F || grad) {
        const dim3 grid(unsigned((n_rows + 127) / 128), n_obj);
#define RFF_CASE(D)                                                                                                      \
    case D:                                                                                                              \
        if (sm > 48 * 1024) {                                                                                            \
            RCU(cudaFuncSetAttribute(rff_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));        \
            RCU(cudaFuncSetAttribute(rff_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));       \
        }                                                                                                                \
        launch_rff<D>(a, grad, grid, sm, nullptr);                                                                       \
        break;
        switch (DP) {
            RFF_CASE(8)
            RFF_CASE(16)
            RFF_CASE(24)
            RFF_CASE(32)
            RFF_CASE(48)
            RFF_CASE(64)
            RFF_CASE(96)
            default:
                if (sm > 48 * 1024) {
                    RCU(cudaFuncSetAttribute(rff_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
                    RCU(cudaFuncSetAttribute(rff_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
                }
                launch_rff<128>(a, grad, grid, sm, nullptr);
                break;
        }
#undef RFF_CASE
        RCU(cudaGetLastError());
    }
    if (hess) {
        rff_hess_kernel<<<dim3(unsigned(n_rows), n_obj), 256, size_t(n_feat) * sizeof(double)>>>(a);
        RCU(cudaGetLastError());
    }
    if (h_F) RCU(cudaMemcpy(h_F, a.F, C * m * sizeof(double), cudaMemcpyDeviceToHost));
    if (grad) RCU(cudaMemcpy(h_dF, a.dF, C * m * d * sizeof(double), cudaMemcpyDeviceToHost));
    if (hess) RCU(cudaMemcpy(h_hF, a.hF, C * m * d * d * sizeof(double), cudaMemcpyDeviceToHost));
    return AOED_OK;
}
