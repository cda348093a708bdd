// Diagonal-block step of the blocked Cholesky factorisation (blocked.cu): one CTA per problem factors the 128 x 128
// diagonal tile  A_kk = L_kk L_kk^T  in shared memory and inverts the factor, so that the panel below it becomes a GEMM
// (L_ik = A_ik L_kk^-T, tgemm_kernel<NT>) and the triangular inverse of the whole matrix can start from X_kk = L_kk^-1.
//
// Inside the tile: blocked right-looking Cholesky on 16 x 16 sub-blocks — the diagonal sub-block is factored by one warp
// in registers with shuffles (the only sequential chain; look-ahead keeps it busy), the panel below is a triangular solve
// with one row per thread, the trailing update runs as DMMA block products; then the eight 16 x 16 inverses at once and
// X = L^-1 in place, block row by block row.  The tile is stored as a full square (row length 132: every DMMA fragment load below is
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
constexpr size_t PD_SMEM = (size_t(PD_N) * PD_LD + PD_NB * FB * (FB + 1) + PD_N) * sizeof(double);

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
    double* Lsm = S + size_t(PD_N) * PD_LD;      // [8][16][17] the factored diagonal sub-blocks
    double* ipv = Lsm + PD_NB * FB * (FB + 1);   // [128] 1 / L_cc
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

    // ---- blocked Cholesky ---------------------------------------------------------------------------------------------
    // The sequential chain belongs to warp 0: it factors the 16 x 16 diagonal sub-block in registers (row r in lane r, pivot
    // column passed around by shuffles).  Nothing else sits on that chain any more:
    //   * the panel below is a triangular SOLVE, one row per thread (112 rows in parallel, 16 short dot products each),
    //     instead of "invert the sub-block, then multiply";
    //   * with look-ahead — after the panel, warp 0 updates only sub-block (kb+1, kb+1) and factors it straight away while
    //     warps 1..15 do the rest of the trailing update;
    //   * the 16 x 16 inverses the in-place inversion below starts from are formed afterwards, all eight at once.
    auto factor_diag = [&](int kb) {
        const int k0 = kb * FB;
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
                    if (lane == c) ipv[k0 + c] = ip;
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
            return;
        }
        if (lane < FB) {
            double* Lk = Lsm + kb * FB * (FB + 1);
#pragma unroll
            for (int c = 0; c < FB; ++c) {
                Lk[lane * (FB + 1) + c] = a[c];
                Kt[int64_t(k0 + lane) * ld + k0 + c] = a[c];  // L sub-block (zeros above the diagonal)
            }
        }
    };
    if (warp == 0) factor_diag(0);
    __syncthreads();
    for (int kb = 0; kb < PD_NB && !s_fail; ++kb) {
        const int k0 = kb * FB;
        // panel: L[i][kb-block] = A[i][kb-block] L_kk^-T, one row per thread by substitution; also kept transposed in the
        // upper sub-blocks [kb][ib]
        {
            const int i = k0 + FB + tid;
            if (i < PD_N) {
                const double* Lk = Lsm + kb * FB * (FB + 1);
                double x[FB];
#pragma unroll
                for (int c = 0; c < FB; ++c) {
                    double sacc = S[i * PD_LD + k0 + c];
#pragma unroll
                    for (int j = 0; j < c; ++j) sacc = fma(-x[j], Lk[c * (FB + 1) + j], sacc);
                    x[c] = sacc * ipv[k0 + c];
                }
#pragma unroll
                for (int c = 0; c < FB; ++c) {
                    S[i * PD_LD + k0 + c] = x[c];
                    S[(k0 + c) * PD_LD + i] = x[c];
                }
            }
        }
        __syncthreads();
        // trailing update: A[ib][jb] -= L[ib][kb] L[jb][kb]^T for kb < jb <= ib.  Pair 0 is the next diagonal sub-block
        {
            const int nt = PD_NB - kb - 1;
            const int npair = nt * (nt + 1) / 2;
            auto do_pair = [&](int pr) {
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
            };
            if (warp == 0) {
                if (npair > 0) {
                    do_pair(0);
                    __syncwarp();
                    factor_diag(kb + 1);
                }
            } else {
                for (int pr = warp; pr < npair; pr += PD_WARPS - 1) do_pair(pr);
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
    // inverses of the eight diagonal sub-blocks, one warp each, one column per lane (forward substitution in registers):
    // the in-place inversion below starts from them
    if (warp < PD_NB && lane < FB) {
        const int k0 = warp * FB, j = lane;
        const double* Lk = Lsm + warp * FB * (FB + 1);
        double x[FB];
#pragma unroll
        for (int i = 0; i < FB; ++i) x[i] = 0.0;
#pragma unroll
        for (int i = 0; i < FB; ++i) {
            if (i < j) continue;
            const double idiag = ipv[k0 + i];
            if (i == j) {
                x[i] = idiag;
            } else {
                double sacc = 0.0;
#pragma unroll
                for (int k = 0; k < FB; ++k)
                    if (k >= j && k < i) sacc = fma(Lk[i * (FB + 1) + k], x[k], sacc);
                x[i] = -sacc * idiag;
            }
        }
#pragma unroll
        for (int i = 0; i < FB; ++i) S[(k0 + i) * PD_LD + k0 + j] = x[i];  // zeros above the diagonal
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
