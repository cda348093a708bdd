// Kernels of the hyper-parameter fit (sm_100a, FP64): kernel matrix, LML gradient contraction, small helpers.
//
//   kmat_kernel       K = c1 * phi(r) + c2 (+ 1e-10 on the diagonal)      (gp.py:24-58 via sklearn kernel algebra,
//                     _gpr.py:582-589)
//   lml_grad_kernel   grad_theta = 1/2 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij/dtheta, fused: the (N,N,p) tensor of
//                     gp.py:63-79 / _gpr.py:627-650 is never materialised.
#pragma once
#include "common.cuh"

namespace aoed {

struct KmatArgs {
    const double* Ts;     // [B][Npad][DP] training points / length scale
    const double* c1c2;   // [B][2]
    double* K;            // [B][Npad][Npad], full symmetric, zero outside N
    int N, Npad, DP;
};

template <int NU, int DP>
__global__ void __launch_bounds__(256) kmat_kernel(KmatArgs p) {
    __shared__ __align__(16) double ti[16][DP];
    __shared__ __align__(16) double tj[16][DP];
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double* __restrict__ Ts = p.Ts + int64_t(b) * p.Npad * DP;
    for (int e = threadIdx.x; e < 16 * DP; e += 256) {
        int r = e / DP, k = e % DP;
        ti[r][k] = (i0 + r < p.N) ? Ts[int64_t(i0 + r) * DP + k] : 0.0;
        tj[r][k] = (j0 + r < p.N) ? Ts[int64_t(j0 + r) * DP + k] : 0.0;
    }
    __syncthreads();
    const int i = i0 + ty, j = j0 + tx;
    if (i >= p.N || j >= p.N) return;
    double r2 = 0.0;
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        double d = ti[ty][k] - tj[tx][k];
        r2 = fma(d, d, r2);
    }
    double phi, gor;
    kernel_profile<NU>(sqrt(r2), r2, phi, gor);
    const double c1 = p.c1c2[2 * b], c2 = p.c1c2[2 * b + 1];
    double v = fma(c1, phi, c2);
    if (i == j) v = (c1 + c2) + kJitter;  // np.fill_diagonal(K, 1) (gp.py:55) then K[diag] += alpha (_gpr.py:589)
    p.K[(int64_t(b) * p.Npad + i) * p.Npad + j] = v;
}

// dphi/dlog l_k = h(r) * D_k, D_k = ((x_ik - x_jk)/l_k)^2   (gp.py:63-79; RBF: sklearn K * D_k)
template <int NU>
__device__ __forceinline__ double h_profile(double r, double r2, double phi) {
    if (NU == 1) return r != 0.0 ? phi / r : 0.0;  // safe_divide + isfinite guard (gp.py:68-70)
    if (NU == 3) return 3.0 * exp(-kSqrt3 * r);
    if (NU == 5) {
        double s = kSqrt5 * r;
        return 5.0 / 3.0 * (s + 1.0) * exp(-s);
    }
    return phi;
}

struct LmlGradArgs {
    const double* Ts;     // [B][Npad][DP]
    const double* c1c2;   // [B][2]
    const double* alpha;  // [B][Npad]
    const double* Kinv;   // [B][Npad][Npad] (row-major lower triangle valid)
    double* partial;      // [B][nBlocks][DP+2]
    int N, Npad, DP, nBlocksX, nBlocksY;
};

// One CTA per 16x16 block of (i,j); per-thread pair; block reduction in shared memory (deterministic), the
// per-block partials are summed by lml_grad_reduce_kernel.  Output order: [c1, l_1..l_DP, c2].
template <int NU, int DP>
__global__ void __launch_bounds__(256) lml_grad_kernel(LmlGradArgs p) {
    __shared__ __align__(16) double ti[16][DP];
    __shared__ __align__(16) double tj[16][DP];
    __shared__ double red[8][DP + 2];
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double* __restrict__ Ts = p.Ts + int64_t(b) * p.Npad * DP;
    for (int e = threadIdx.x; e < 16 * DP; e += 256) {
        int r = e / DP, k = e % DP;
        ti[r][k] = (i0 + r < p.N) ? Ts[int64_t(i0 + r) * DP + k] : 0.0;
        tj[r][k] = (j0 + r < p.N) ? Ts[int64_t(j0 + r) * DP + k] : 0.0;
    }
    __syncthreads();
    const int i = i0 + ty, j = j0 + tx;
    double g[DP + 2];
#pragma unroll
    for (int k = 0; k < DP + 2; ++k) g[k] = 0.0;
    if (i < p.N && j < p.N) {
        const double c1 = p.c1c2[2 * b], c2 = p.c1c2[2 * b + 1];
        const double* alpha = p.alpha + int64_t(b) * p.Npad;
        // K^-1 comes from potri: only the row-major LOWER triangle is valid, read it symmetrically
        const double w = alpha[i] * alpha[j] - p.Kinv[(int64_t(b) * p.Npad + max(i, j)) * p.Npad + min(i, j)];
        double dk[DP];
        double r2 = 0.0;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            double d = ti[ty][k] - tj[tx][k];
            dk[k] = d * d;
            r2 += dk[k];
        }
        const double r = sqrt(r2);
        double phi, gor;
        kernel_profile<NU>(r, r2, phi, gor);
        if (i == j) phi = 1.0;
        g[0] = w * c1 * phi;
        const double wh = w * c1 * h_profile<NU>(r, r2, phi);
#pragma unroll
        for (int k = 0; k < DP; ++k) g[1 + k] = wh * dk[k];
        g[DP + 1] = w * c2;
    }
    // warp reduce, then across the 8 warps
#pragma unroll
    for (int k = 0; k < DP + 2; ++k) {
        double v = g[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < DP + 2) {
        double s = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
        const int64_t blk = int64_t(blockIdx.y) * p.nBlocksX + blockIdx.x;
        p.partial[(int64_t(b) * p.nBlocksX * p.nBlocksY + blk) * (DP + 2) + threadIdx.x] = s;
    }
}

// res[1 + k] = 0.5 * sum_blocks partial (k in the theta order [c1, l_1..l_d, c2]); zeros when the factorisation failed.
// One CTA per component k (deterministic tree reduction).
static __global__ void __launch_bounds__(256) lml_grad_reduce_kernel(const double* __restrict__ partial, int nBlocks, int DP, int d,
                                                              const int* info, double* res) {
    __shared__ double red[8];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < nBlocks; i += 256) s += partial[int64_t(i) * (DP + 2) + k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x != 0) return;
    s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    s = (*info != 0) ? 0.0 : 0.5 * s;
    if (k == 0) res[1] = s;
    else if (k == DP + 1) res[1 + d + 1] = s;
    else if (k - 1 < d) res[1 + k] = s;
}

// y_out[i] = sum_j A[i][j] x[j] over the stored triangle of a row-major matrix: LOWER (j <= i) or UPPER (j >= i).
// One warp per row; used for z = L^-1 y and alpha = L^-T z of the blocked LML path.
template <bool UPPER>
__global__ void __launch_bounds__(256) tri_matvec_kernel(const double* __restrict__ A, const double* __restrict__ x, int N, int ld,
                                                         double* __restrict__ y_out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const double* a = A + int64_t(row) * ld;
    const int j0 = UPPER ? row : 0, j1 = UPPER ? N : row + 1;
    double s = 0.0;
    for (int j = j0 + lane; j < j1; j += 32) s = fma(a[j], x[j], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y_out[row] = s;
}

// Ts = X / l (padded to DP), c1c2 = exp(theta[0]), exp(theta[d+1]) for one hyper-parameter vector (device-side, no host sync)
static __global__ void theta_prep_kernel(const double* __restrict__ X, const double* __restrict__ theta, int N, int Npad, int d,
                                  int DP, double* __restrict__ Ts, double* __restrict__ c1c2) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) {
        c1c2[0] = exp(theta[0]);
        c1c2[1] = exp(theta[d + 1]);
    }
    if (idx >= Npad * DP) return;
    const int i = idx / DP, k = idx % DP;
    Ts[idx] = (i < N && k < d) ? X[int64_t(i) * d + k] / exp(theta[1 + k]) : 0.0;
}

// res[0] = -1/2 y.alpha - sum log L_ii - N/2 log 2 pi   (_gpr.py:601-616); -inf when potrf reported a non-PD matrix
// quad_a . quad_b is y . alpha (sklearn's form) or z . z with z = L^-1 y (same number, better conditioned)
static __global__ void __launch_bounds__(256) lml_value_kernel(const double* __restrict__ y, const double* __restrict__ alpha,
                                                        const double* __restrict__ Lfac, int N, int Npad, const int* info,
                                                        double* res) {
    __shared__ double red[2][8];
    double q = 0.0, ld = 0.0;
    for (int i = threadIdx.x; i < N; i += 256) {
        q = fma(y[i], alpha[i], q);
        ld += log(Lfac[int64_t(i) * Npad + i]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        q += __shfl_xor_sync(0xffffffffu, q, off);
        ld += __shfl_xor_sync(0xffffffffu, ld, off);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = q;
        red[1][threadIdx.x >> 5] = ld;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double qq = 0.0, ll = 0.0;
        for (int w = 0; w < 8; ++w) {
            qq += red[0][w];
            ll += red[1][w];
        }
        res[0] = (*info != 0) ? -INFINITY : -0.5 * qq - ll - 0.5 * N * 1.8378770664093453;
    }
}

// out[j][i] = in[i][j] for a square Npad matrix (batched via blockIdx.z)
static __global__ void transpose_square_kernel(const double* in, double* out, int n) {
    __shared__ double tile[32][33];
    const double* src = in + int64_t(blockIdx.z) * n * n;
    double* dst = out + int64_t(blockIdx.z) * n * n;
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
    for (int r = 0; r < 32; r += 8)
        if (x < n && y + r < n) tile[threadIdx.y + r][threadIdx.x] = src[int64_t(y + r) * n + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x;
    y = blockIdx.x * 32 + threadIdx.y;
    for (int r = 0; r < 32; r += 8)
        if (x < n && y + r < n) dst[int64_t(y + r) * n + x] = tile[threadIdx.x][threadIdx.y + r];
}

// keeps the lower triangle (row-major, rows/cols < N) and zeroes everything else
static __global__ void keep_lower_kernel(double* A, int N, int Npad) {
    double* M = A + int64_t(blockIdx.z) * Npad * Npad;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= Npad) return;
    if (i >= N || j >= N || j > i) M[int64_t(i) * Npad + j] = 0.0;
}

static __global__ void set_identity_kernel(double* A, int N, int Npad) {
    double* M = A + int64_t(blockIdx.z) * Npad * Npad;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= Npad) return;
    M[int64_t(i) * Npad + j] = (i == j && i < N) ? 1.0 : 0.0;
}

}  // namespace aoed
