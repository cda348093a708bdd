// Random-Fourier-feature sample evaluation kernels (see rff.cu for the contract), shared with the device NSGA-II.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aoed {

struct RffArgs {
    const double* X;       // [rows][d] normalised candidates
    const double* W;       // [m][M][d]
    const double* b;       // [m][M]
    const double* theta;   // [m][M]
    const double* factor;  // [m]  sqrt(2 sf2 / M)
    const double* y_mean;  // [m]
    const double* y_scale; // [m]
    double *F, *dF, *hF;   // (rows, m), (rows, m, d), (rows, m, d, d); any may be null
    int64_t rows;
    int d, m, M;
};

template <int DP, bool GRAD>
__global__ void __launch_bounds__(128) rff_kernel(RffArgs p) {
    extern __shared__ double rff_smem[];  // [M][DP] W (zero padded), [M] b, [M] theta
    const int o = blockIdx.y, d = p.d, M = p.M;
    double* sW = rff_smem;
    double* sb = sW + size_t(M) * DP;
    double* st = sb + M;
    for (int i = threadIdx.x; i < M * DP; i += 128) {
        const int j = i / DP, k = i % DP;
        sW[i] = k < d ? p.W[(int64_t(o) * M + j) * d + k] : 0.0;
    }
    for (int j = threadIdx.x; j < M; j += 128) {
        sb[j] = p.b[int64_t(o) * M + j];
        st[j] = p.theta[int64_t(o) * M + j];
    }
    __syncthreads();
    const int64_t c = blockIdx.x * int64_t(128) + threadIdx.x;
    if (c >= p.rows) return;
    double x[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) x[k] = k < d ? p.X[c * d + k] : 0.0;
    double f = 0.0, g[GRAD ? DP : 1];
    if (GRAD) {
#pragma unroll
        for (int k = 0; k < DP; ++k) g[k] = 0.0;
    }
    for (int j = 0; j < M; ++j) {
        const double* w = sW + size_t(j) * DP;
        double t = sb[j];
#pragma unroll
        for (int k = 0; k < DP; ++k) t = fma(w[k], x[k], t);
        double sn, cs;
        sincos(t, &sn, &cs);
        const double th = st[j];
        f = fma(th, cs, f);
        if (GRAD) {
            const double q = th * sn;
#pragma unroll
            for (int k = 0; k < DP; ++k) g[k] = fma(q, w[k], g[k]);
        }
    }
    const double fac = p.factor[o], ys = p.y_scale[o];
    if (p.F) p.F[c * p.m + o] = fma(fac * f, ys, p.y_mean[o]);  // normalization.undo (ts.py:83)
    if (GRAD && p.dF) {
#pragma unroll
        for (int k = 0; k < DP; ++k)
            if (k < d) p.dF[(c * p.m + o) * d + k] = -fac * g[k] * ys;  // ts.py:74, rescale ts.py:84
    }
}

static __global__ void __launch_bounds__(256) rff_hess_kernel(RffArgs p) {
    extern __shared__ double hs_smem[];  // [M] theta_j cos(t_j)
    const int64_t c = blockIdx.x;
    const int o = blockIdx.y, d = p.d, M = p.M;
    const double* W = p.W + int64_t(o) * M * d;
    for (int j = threadIdx.x; j < M; j += 256) {
        double t = p.b[int64_t(o) * M + j];
        for (int k = 0; k < d; ++k) t = fma(W[int64_t(j) * d + k], p.X[c * d + k], t);
        hs_smem[j] = p.theta[int64_t(o) * M + j] * cos(t);
    }
    __syncthreads();
    const double fac = p.factor[o];
    for (int pr = threadIdx.x; pr < d * d; pr += 256) {
        const int i = pr / d, k = pr % d;
        double s = 0.0;
        for (int j = 0; j < M; ++j) s = fma(hs_smem[j] * W[int64_t(j) * d + i], W[int64_t(j) * d + k], s);
        p.hF[((c * p.m + o) * d + i) * d + k] = -fac * s;  // ts.py:77 (not rescaled: literal)
    }
}


inline int rff_dp(int n_var) {
    return n_var <= 8 ? 8 : n_var <= 16 ? 16 : n_var <= 24 ? 24 : n_var <= 32 ? 32 : n_var <= 48 ? 48 : n_var <= 64 ? 64 : n_var <= 96 ? 96 : 128;
}
inline size_t rff_smem_bytes(int n_feat, int DP) { return (size_t(n_feat) * DP + 2 * size_t(n_feat)) * sizeof(double); }

// value-only launch on device buffers (used by the device-resident NSGA-II)
inline cudaError_t rff_launch_value(const RffArgs& a, cudaStream_t s) {
    const int DP = rff_dp(a.d);
    const size_t sm = rff_smem_bytes(a.M, DP);
    const dim3 grid(unsigned((a.rows + 127) / 128), a.m);
    cudaError_t e = cudaSuccess;
    switch (DP) {
        case 8:
            if (sm > 48 * 1024) e = cudaFuncSetAttribute(rff_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));
            rff_kernel<8, false><<<grid, 128, sm, s>>>(a);
            break;
        case 16:
            if (sm > 48 * 1024) e = cudaFuncSetAttribute(rff_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));
            rff_kernel<16, false><<<grid, 128, sm, s>>>(a);
            break;
        case 24:
            if (sm > 48 * 1024) e = cudaFuncSetAttribute(rff_kernel<24, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));
            rff_kernel<24, false><<<grid, 128, sm, s>>>(a);
            break;
        case 32:
            if (sm > 48 * 1024) e = cudaFuncSetAttribute(rff_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));
            rff_kernel<32, false><<<grid, 128, sm, s>>>(a);
            break;
        case 48:
            if (sm > 48 * 1024) e = cudaFuncSetAttribute(rff_kernel<48, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));
            rff_kernel<48, false><<<grid, 128, sm, s>>>(a);
            break;
        default:  // the device-resident NSGA-II loop, the only caller, takes n_var <= 64
            if (sm > 48 * 1024) e = cudaFuncSetAttribute(rff_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm));
            rff_kernel<64, false><<<grid, 128, sm, s>>>(a);
            break;
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace aoed
