// Fused log-marginal-likelihood + gradient for a BATCH of small GP problems (sm_100a, FP64): one CTA per
// (objective, restart) problem, the whole kernel matrix resident in shared memory as a packed lower triangle.
//
// Replaces, per problem, sklearn GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True)
// (_gpr.py:541-655, reached from autooed/mobo/surrogate_model/gp.py:139): kernel matrix + Cholesky + alpha +
// K^-1 + the (alpha alpha^T - K^-1) : dK/dtheta contraction, for every L-BFGS-B problem of the batch in one launch.
// The (N,N,d) gradient tensor of gp.py:63-79 is never formed.  N(N+1)/2 doubles must fit in shared memory
// (N <= 232 with the 227 KB opt-in) — the MOBO regime of the reference (N = 20 ... ~220); larger problems take the
// blocked path in gp.cu.
//
// Phases (all in shared memory, CTA-wide barriers in between):
//   1  K = c1 phi(r) + c2 (+ jitter on the diagonal), lower triangle only
//   2  in-place right-looking Cholesky  K = L L^T          (a non-positive pivot => lml = -inf, grad = 0, _gpr.py:589-593)
//   3  in-place inverse of L (right-looking rank-1 updates, no reductions)
//   4  z = L^-1 y, alpha = L^-T z, lml = -1/2 z.z - sum log L_ii - N/2 log 2 pi
//   5  grad_k = 1/2 sum_ij (alpha_i alpha_j - (L^-T L^-1)_ij) dK_ij/dtheta_k with K^-1_ij formed on the fly
#pragma once
#include "common.cuh"
#include "fit.cuh"

namespace aoed {

constexpr int FS_THREADS = 512;
constexpr int FS_WARPS = FS_THREADS / 32;
constexpr int FS_MAX_N = 232;

struct LmlSmallArgs {
    const double* X;       // [N][d] normalised training inputs
    const double* XT;      // [d][Npad] the same, dimension-major and zero padded (coalesced 4 x 4 pair blocks in lml_blk_kernel)
    const double* Y;       // [m][Npad] normalised targets, objective-major
    const double* theta;   // [B][d+2]
    const int* obj;        // [B]
    double* res;           // [B][1 + d + 2]: lml, grad[c1, l_1..l_d, c2]
    int N, Npad, d, want_grad;
    long long* clk;        // optional [B][8] phase timestamps (clock64) for tuning, or nullptr
};

inline size_t lml_small_smem(int N, int DP) {
    return (size_t(N) * (N + 1) / 2 + 4 * size_t(N) + 2 * DP + FS_WARPS * (DP + 2) + 8) * sizeof(double);
}

__device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

template <int NU, int DP>
__global__ void __launch_bounds__(FS_THREADS, 1) lml_small_kernel(LmlSmallArgs p) {
    extern __shared__ __align__(16) double fs_smem[];
    const int N = p.N, d = p.d;
    double* Lp = fs_smem;                       // packed lower triangle: (i, j<=i) at tri(i) + j
    double* col = Lp + size_t(N) * (N + 1) / 2;  // [N] column scratch
    double* yv = col + N;                        // [N]
    double* zv = yv + N;                         // [N]
    double* al = zv + N;                         // [N]
    double* iell = al + N;                       // [DP] 1 / length scale (0 for padded dims)
    double* misc = iell + DP;                    // [DP] scratch + flags
    double* red = misc + DP + 8;                 // [FS_WARPS][DP + 2]
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const double* __restrict__ th = p.theta + int64_t(b) * (d + 2);
    const double c1 = exp(th[0]), c2 = exp(th[d + 1]);
    if (tid < DP) iell[tid] = (tid < d) ? exp(-th[1 + tid]) : 0.0;
    if (tid == 0) s_fail = 0;
    const double* __restrict__ y = p.Y + int64_t(p.obj[b]) * p.Npad;
    for (int i = tid; i < N; i += FS_THREADS) yv[i] = y[i];
    __syncthreads();

#define FS_STAMP(idx)                                                  \
    if (p.clk != nullptr && tid == 0) p.clk[int64_t(b) * 8 + (idx)] = clock64();
    FS_STAMP(0)
    // ---- 1: kernel matrix -------------------------------------------------------------------------------------
    for (int i = warp; i < N; i += FS_WARPS) {
        const double* xi = p.X + int64_t(i) * d;
        for (int j = lane; j <= i; j += 32) {
            const double* xj = p.X + int64_t(j) * d;
            double r2 = 0.0;
            for (int k = 0; k < d; ++k) {
                const double t = (xi[k] - xj[k]) * iell[k];
                r2 = fma(t, t, r2);
            }
            double phi, gor;
            kernel_profile<NU>(sqrt(r2), r2, phi, gor);
            Lp[tri(i) + j] = (i == j) ? (c1 + c2) + kJitter : fma(c1, phi, c2);  // gp.py:55, _gpr.py:589
        }
    }
    __syncthreads();

    FS_STAMP(1)
    // ---- 2: Cholesky -------------------------------------------------------------------------------------------
    for (int k = 0; k < N; ++k) {
        const double dkk = Lp[tri(k) + k];
        if (!(dkk > 0.0)) {  // also catches NaN
            if (tid == 0) s_fail = 1;
            break;  // uniform: every thread reads the same value
        }
        // every thread needs 1/pivot: rsqrt (one MUFU + Newton steps) instead of sqrt + divide; log|L| is summed after the loop
        const double ipiv = rsqrt(dkk);
        const double piv = dkk * ipiv;
        __syncthreads();  // everyone has read the pivot
        for (int i = k + 1 + tid; i < N; i += FS_THREADS) {
            const double v = Lp[tri(i) + k] * ipiv;
            Lp[tri(i) + k] = v;
            col[i] = v;
        }
        if (tid == 0) Lp[tri(k) + k] = piv;
        __syncthreads();
        for (int i = k + 1 + warp; i < N; i += FS_WARPS) {
            const double ci = col[i];
            double* row = Lp + tri(i);
            for (int j = k + 1 + lane; j <= i; j += 32) row[j] = fma(-ci, col[j], row[j]);
        }
        __syncthreads();
    }
    __syncthreads();
    double* out = p.res + int64_t(b) * (d + 3);
    double logdet = 0.0;  // complete in every lane of warp 0 only (the one consumer)
    if (warp == 0 && !s_fail) {
        for (int k = lane; k < N; k += 32) logdet += log(Lp[tri(k) + k]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) logdet += __shfl_xor_sync(0xffffffffu, logdet, off);
    }
    if (s_fail) {
        if (tid == 0) out[0] = -INFINITY;
        if (tid < d + 2) out[1 + tid] = 0.0;
        return;
    }

    FS_STAMP(2)
    // ---- 3: L <- L^-1 in place ---------------------------------------------------------------------------------
    // Right-looking: storage holds R = I - (partial products) below/left of the pivot row; step k finalises row k of
    // X = L^-1 (X[k][j] = R[k][j] / L_kk) and applies the rank-1 update R[i][j] -= L[i][k] X[k][j] to the rectangle
    // i > k, j <= k.  Entry (i, k) turns from L[i][k] into R[i][k] at step k, so the inversion is in place, and every
    // update is an independent fused multiply-add (no reductions).
    double* rowb = zv;  // [N] scratch: row k of X (zv is not used before phase 4)
    for (int k = 0; k < N; ++k) {
        const double ikk = 1.0 / Lp[tri(k) + k];
        __syncthreads();  // everyone has read the pivot; the previous step's updates are complete
        for (int i = k + 1 + tid; i < N; i += FS_THREADS) col[i] = Lp[tri(i) + k];
        for (int j = tid; j <= k; j += FS_THREADS) {
            const double x = (j == k) ? ikk : Lp[tri(k) + j] * ikk;
            Lp[tri(k) + j] = x;
            rowb[j] = x;
        }
        __syncthreads();
        for (int i = k + 1 + warp; i < N; i += FS_WARPS) {
            const double ci = col[i];
            double* row = Lp + tri(i);
            for (int j = lane; j <= k; j += 32) row[j] = fma(-ci, rowb[j], (j == k) ? 0.0 : row[j]);
        }
    }
    __syncthreads();

    FS_STAMP(3)
    // ---- 4: z = L^-1 y, alpha = L^-T z ---------------------------------------------------------------------------
    for (int i = warp; i < N; i += FS_WARPS) {
        const double* row = Lp + tri(i);
        double s = 0.0;
        for (int k = lane; k <= i; k += 32) s = fma(row[k], yv[k], s);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) zv[i] = s;
    }
    __syncthreads();
    for (int j = warp; j < N; j += FS_WARPS) {
        double s = 0.0;
        for (int i = j + lane; i < N; i += 32) s = fma(Lp[tri(i) + j], zv[i], s);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) al[j] = s;
    }
    if (warp == 0) {
        double q = 0.0;
        for (int i = lane; i < N; i += 32) q = fma(zv[i], zv[i], q);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
        if (lane == 0) out[0] = -0.5 * q - logdet - 0.5 * N * 1.8378770664093453;  // log(2 pi)   (_gpr.py:601-616)
    }
    __syncthreads();
    FS_STAMP(4)
    if (!p.want_grad) return;

    // ---- 5: gradient ---------------------------------------------------------------------------------------------
    double g[DP + 2];
#pragma unroll
    for (int k = 0; k < DP + 2; ++k) g[k] = 0.0;
    for (int i = warp; i < N; i += FS_WARPS) {
        const double* xi = p.X + int64_t(i) * d;
        const double ai = al[i];
        for (int j = lane; j <= i; j += 32) {
            // K^-1_ij = sum_{k >= i} Linv[k][i] * Linv[k][j]
            double kinv = 0.0;
            for (int k = i; k < N; ++k) kinv = fma(Lp[tri(k) + i], Lp[tri(k) + j], kinv);
            double w = ai * al[j] - kinv;  // _gpr.py:627-633
            if (i != j) w *= 2.0;          // symmetric pair
            const double* xj = p.X + int64_t(j) * d;
            double r2 = 0.0;
            for (int k = 0; k < d; ++k) {
                const double t = (xi[k] - xj[k]) * iell[k];
                r2 = fma(t, t, r2);
            }
            const double r = sqrt(r2);
            double phi, gor;
            kernel_profile<NU>(r, r2, phi, gor);
            if (i == j) phi = 1.0;
            g[0] = fma(w * c1, phi, g[0]);
            const double wh = w * c1 * h_profile<NU>(r, r2, phi);
#pragma unroll
            for (int k = 0; k < DP; ++k) {  // D_k recomputed instead of held in registers (gp.py:63-79)
                const double t = (k < d) ? (xi[k] - xj[k]) * iell[k] : 0.0;
                g[1 + k] = fma(wh, t * t, g[1 + k]);
            }
            g[DP + 1] = fma(w, c2, g[DP + 1]);
        }
    }
#pragma unroll
    for (int k = 0; k < DP + 2; ++k) {
        double v = g[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp * (DP + 2) + k] = v;
    }
    __syncthreads();
    if (tid < DP + 2) {
        double s = 0.0;
        for (int w8 = 0; w8 < FS_WARPS; ++w8) s += red[w8 * (DP + 2) + tid];
        s *= 0.5;
        if (tid == 0) out[1] = s;
        else if (tid == DP + 1) out[1 + d + 1] = s;
        else if (tid - 1 < d) out[1 + tid] = s;
    }
    FS_STAMP(5)
#undef FS_STAMP
}

}  // namespace aoed
