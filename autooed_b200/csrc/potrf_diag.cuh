// Diagonal-block step of the blocked Cholesky factorisation (blocked.cu): one CTA per problem factors the 128 x 128
// diagonal tile  A_kk = L_kk L_kk^T  in shared memory and inverts the factor, so that the panel below it becomes a GEMM
// (L_ik = A_ik L_kk^-T, tgemm_kernel<NT>) and the triangular inverse of the whole matrix can start from X_kk = L_kk^-1.
//
// Inside the tile the algorithm is the blocked right-looking one of lml_blk_kernel (fit_small2.cuh) on 16 x 16
// sub-blocks: the diagonal sub-block is factored by one warp in registers with shuffles and inverted by forward
// substitution, panel and trailing update run as DMMA block products on all 16 warps; then X = L^-1 in place, block
// row by block row.  The tile is stored as a full square (row length 132: every DMMA fragment load below is
// bank-conflict free), the strictly upper sub-blocks hold transposed copies of the panels so that the in-place
// inverse reads L[ib][kb] from there while the lower sub-block is being overwritten.
//
// Outputs: L_kk (lower, zeros above the diagonal) back into the matrix, L_kk^-1 (lower, zeros above) into the same tile
// of the inverse.  A non-positive pivot sets info[problem] = 1 + global pivot index (first failure wins) and leaves
// identity tiles behind so that everything downstream stays finite; the caller reports lml = -inf for that problem.
#pragma once
#include "common.cuh"
#include "fit_small2.cuh"

namespace aoed {

constexpr int PD_N = 128, PD_LD = 132, PD_NB = PD_N / FB;
constexpr int PD_THREADS = 512, PD_WARPS = PD_THREADS / 32;
constexpr size_t PD_SMEM = (size_t(PD_N) * PD_LD + 2 * FB * (FB + 1)) * sizeof(double);

struct PotrfDiagArgs {
    double* K;       // [batch][Npad][ld]: tile (k, k) is read, L_kk written back
    double* Xinv;    // [batch][Npad][ld]: tile (k, k) receives L_kk^-1
    int64_t stride;  // elements between problems (both matrices)
    int ld;
    int k;           // tile index
    int* info;       // [batch]
};

__global__ void __launch_bounds__(PD_THREADS, 1) potrf_diag_kernel(PotrfDiagArgs p) {
    extern __shared__ __align__(16) double pd_smem[];
    double* S = pd_smem;                     // [128][132]
    double* Dsm = S + size_t(PD_N) * PD_LD;  // [16][17] diagonal sub-block
    double* Dinv = Dsm + FB * (FB + 1);      // [16][17] its inverse
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int b = blockIdx.x;
    const int64_t off = int64_t(b) * p.stride + int64_t(p.k) * PD_N * p.ld + int64_t(p.k) * PD_N;
    double* __restrict__ Kt = p.K + off;
    double* __restrict__ Xt = p.Xinv + off;
    const int ld = p.ld;

    auto write_identity = [&]() {
        for (int e = tid; e < PD_N * PD_N; e += PD_THREADS) {
            const int r = e >> 7, c = e & 127;
            const double v = (r == c) ? 1.0 : 0.0;
            Kt[int64_t(r) * ld + c] = v;
            Xt[int64_t(r) * ld + c] = v;
        }
    };
    if (p.info[b] != 0) {  // an earlier tile of this problem failed
        write_identity();
        return;
    }
    if (tid == 0) s_fail = 0;
    for (int e = tid; e < PD_N * PD_N / 2; e += PD_THREADS) {
        const int r = e >> 6, c2 = (e & 63) * 2;
        *reinterpret_cast<double2*>(S + r * PD_LD + c2) = *reinterpret_cast<const double2*>(Kt + int64_t(r) * ld + c2);
    }
    __syncthreads();

    // ---- blocked Cholesky; the diagonal sub-blocks end up holding their inverses ------------------------------------
    for (int kb = 0; kb < PD_NB; ++kb) {
        const int k0 = kb * FB;
        if (warp == 0) {
            // row r of the sub-block in the registers of lane r (lanes 16..31 mirror 0..15), pivot column by shuffles
            const int r = lane & 15;
            double a[FB];
#pragma unroll
            for (int c = 0; c < FB; ++c) a[c] = (c <= r) ? S[(k0 + r) * PD_LD + k0 + c] : 0.0;
            int bad = 0;
#pragma unroll
            for (int c = 0; c < FB; ++c) {
                if (!bad) {
                    const double dcc = __shfl_sync(0xffffffffu, a[c], c);
                    if (!(dcc > 0.0)) {
                        bad = c + 1;  // the same value in every lane: uniform
                    } else {
                        const double ip = rsqrt(dcc);
                        a[c] = (r == c) ? dcc * ip : a[c] * ip;
                        if (lane == c) Dinv[c * (FB + 1) + c] = ip;
#pragma unroll
                        for (int cc = c + 1; cc < FB; ++cc) {
                            const double lcc = __shfl_sync(0xffffffffu, a[c], cc);
                            if (cc <= r) a[cc] = fma(-a[c], lcc, a[cc]);
                        }
                    }
                }
            }
            if (bad) {
                if (lane == 0) s_fail = p.k * PD_N + k0 + bad;
            } else {
                if (lane < FB) {
#pragma unroll
                    for (int c = 0; c < FB; ++c) {
                        Dsm[lane * (FB + 1) + c] = a[c];
                        Kt[int64_t(k0 + lane) * ld + k0 + c] = a[c];  // L sub-block (zeros above the diagonal)
                    }
                }
                __syncwarp();
                // inverse of the lower-triangular sub-block, one column per lane (forward substitution in registers)
                if (lane < FB) {
                    const int j = lane;
                    double x[FB];
#pragma unroll
                    for (int i = 0; i < FB; ++i) x[i] = 0.0;
#pragma unroll
                    for (int i = 0; i < FB; ++i) {
                        if (i < j) continue;
                        const double idiag = Dinv[i * (FB + 1) + i];
                        if (i == j) {
                            x[i] = idiag;
                        } else {
                            double s = 0.0;
#pragma unroll
                            for (int k = 0; k < FB; ++k)
                                if (k >= j && k < i) s = fma(Dsm[i * (FB + 1) + k], x[k], s);
                            x[i] = -s * idiag;
                        }
                    }
                    __syncwarp(0xffff);
#pragma unroll
                    for (int i = 0; i < FB; ++i) Dinv[i * (FB + 1) + j] = x[i];  // upper part = 0
                }
                __syncwarp();
                for (int e = lane; e < FB * FB; e += 32) {
                    const int rr = e / FB, c = e % FB;
                    S[(k0 + rr) * PD_LD + k0 + c] = Dinv[rr * (FB + 1) + c];
                }
            }
        }
        __syncthreads();
        if (s_fail) break;
        // panel: L[ib][kb] = A[ib][kb] * Dinv^T; also kept transposed in the upper sub-block [kb][ib]
        for (int ib = kb + 1 + warp; ib < PD_NB; ib += PD_WARPS) {
            const int i0 = ib * FB;
            BlkAcc acc;
            acc.zero();
            blk_mma(acc, [&](int r, int k) { return S[(i0 + r) * PD_LD + k0 + k]; },
                    [&](int k, int n) { return Dinv[n * (FB + 1) + k]; }, g, q);
            __syncwarp();
            blk_foreach(acc, [&](int r, int c, double& v) {
                S[(i0 + r) * PD_LD + k0 + c] = v;
                S[(k0 + c) * PD_LD + i0 + r] = v;
            }, g, q);
        }
        __syncthreads();
        // trailing update: A[ib][jb] -= L[ib][kb] L[jb][kb]^T for kb < jb <= ib
        {
            const int nt = PD_NB - kb - 1;
            const int npair = nt * (nt + 1) / 2;
            for (int pr = warp; pr < npair; pr += PD_WARPS) {
                int bi = 0;
                while ((bi + 1) * (bi + 2) / 2 <= pr) ++bi;
                const int bj = pr - bi * (bi + 1) / 2;
                const int i0 = (kb + 1 + bi) * FB, j0 = (kb + 1 + bj) * FB;
                BlkAcc acc;
                acc.zero();
                blk_mma(acc, [&](int r, int k) { return S[(i0 + r) * PD_LD + k0 + k]; },
                        [&](int k, int n) { return S[(j0 + n) * PD_LD + k0 + k]; }, g, q);
                blk_foreach(acc, [&](int r, int c, double& v) {
                    if (j0 + c <= i0 + r) S[(i0 + r) * PD_LD + j0 + c] -= v;
                }, g, q);
            }
        }
        __syncthreads();
    }
    if (s_fail) {
        if (tid == 0) atomicCAS(p.info + b, 0, s_fail);
        write_identity();
        return;
    }
    // L_kk: strictly lower sub-blocks from shared memory, zeros above (the diagonal sub-blocks were written above)
    for (int e = tid; e < PD_N * PD_N; e += PD_THREADS) {
        const int r = e >> 7, c = e & 127;
        if ((r >> 4) == (c >> 4)) continue;
        Kt[int64_t(r) * ld + c] = (c < r) ? S[r * PD_LD + c] : 0.0;
    }
    __syncthreads();

    // ---- X = L^-1 in place, block rows top to bottom --------------------------------------------------------------------
    // before step kb: block rows < kb hold X; rows >= kb hold R = -(partial products) in columns < kb, L in columns >= kb
    // (strictly lower sub-blocks; their transposes sit in the upper sub-blocks) and L_kk^-1 on the diagonal
    for (int kb = 0; kb < PD_NB; ++kb) {
        const int k0 = kb * FB;
        for (int jb = warp; jb < kb; jb += PD_WARPS) {  // X[kb][jb] = Dinv_kb * R[kb][jb]
            const int j0 = jb * FB;
            BlkAcc acc;
            acc.zero();
            blk_mma(acc, [&](int r, int k) { return S[(k0 + r) * PD_LD + k0 + k]; },
                    [&](int k, int n) { return S[(k0 + k) * PD_LD + j0 + n]; }, g, q);
            __syncwarp();
            blk_foreach(acc, [&](int r, int c, double& v) { S[(k0 + r) * PD_LD + j0 + c] = v; }, g, q);
        }
        __syncthreads();
        {  // R[ib][jb] = (jb == kb ? 0 : R[ib][jb]) - L[ib][kb] * X[kb][jb]   for ib > kb, jb <= kb
            const int nrow = PD_NB - kb - 1, ncol = kb + 1;
            for (int pr = warp; pr < nrow * ncol; pr += PD_WARPS) {
                const int ib = kb + 1 + pr / ncol, jb = pr % ncol;
                const int i0 = ib * FB, j0 = jb * FB;
                BlkAcc acc;
                acc.zero();
                blk_mma(acc, [&](int r, int k) { return S[(k0 + k) * PD_LD + i0 + r]; },
                        [&](int k, int n) { return S[(k0 + k) * PD_LD + j0 + n]; }, g, q);
                blk_foreach(acc, [&](int r, int c, double& v) {
                    double* dst = S + (i0 + r) * PD_LD + j0 + c;
                    *dst = (jb == kb ? 0.0 : *dst) - v;
                }, g, q);
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < PD_N * PD_N; e += PD_THREADS) {
        const int r = e >> 7, c = e & 127;
        Xt[int64_t(r) * ld + c] = (c <= r) ? S[r * PD_LD + c] : 0.0;
    }
}

}  // namespace aoed
