// Blocked version of the one-CTA-per-problem log-marginal-likelihood + gradient kernel (sm_100a, FP64 DMMA).
//
// Same contract and shared-memory layout (packed lower triangle) as lml_small_kernel (fit_small.cuh); the three
// O(N^3) phases are restructured into 16 x 16 blocks whose products run on the FP64 tensor pipe (mma.sync m8n8k4,
// SASS DMMA) instead of element-wise rank-1 sweeps with a CTA barrier per column:
//
//   Cholesky   right-looking blocked: diagonal block factored + inverted by one warp, panel  L21 = A21 L11^-T  and
//              trailing update  A22 -= L21 L21^T  as block GEMMs spread over all 16 warps
//   inverse    X = L^-1 in place, block row at a time:  X[k][j] = L11^-1 R[k][j],  R[i][j] -= L[i][k] X[k][j]
//   K^-1       G = X^T X in place, block row at a time: G[i][j] += X[t][i]^T X[t][j]
//   gradient   element-wise contraction of (alpha alpha^T - K^-1) with dK/dtheta (no reductions inside the pair loop)
//
// Barriers: 3 per block column (13 block columns at N = 200) instead of 3 per column.
#pragma once
#include "common.cuh"
#include "fit.cuh"
#include "fit_small.cuh"

namespace aoed {

constexpr int FB = 16;       // block size
constexpr int FBP = FB + 1;  // padded row length of the panel buffer (bank-conflict free fragment loads)

// packed triangle + ONE panel buffer + small vectors; the caller falls back to lml_small_kernel when this exceeds 227 KB
inline size_t lml_blk_smem(int N, int DP) {
    const size_t Nr = size_t((N + FB - 1) / FB) * FB;
    return (size_t(N) * (N + 1) / 2 + 4 * size_t(N) + FBP * Nr + 2 * FB * (FB + 1) + 2 * DP + FS_WARPS * (DP + 2) + 16) *
           sizeof(double);
}

// accumulators of one 16 x 16 block: 2 x 2 DMMA tiles; lane (g, q) owns C[8 mi + g][8 ni + 2 q + e]
struct BlkAcc {
    double c[2][2][2];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) c[a][b][0] = c[a][b][1] = 0.0;
    }
};

// acc += A (16 x 16) * B (16 x 16); fa(r, k) and fb(k, n) return operand elements
template <class FA, class FB_>
__device__ __forceinline__ void blk_mma(BlkAcc& acc, FA fa, FB_ fb, int g, int q) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int k = 4 * s + q;
        const double a0 = fa(g, k), a1 = fa(8 + g, k);
        const double b0 = fb(k, g), b1 = fb(k, 8 + g);
        dmma884(acc.c[0][0][0], acc.c[0][0][1], a0, b0);
        dmma884(acc.c[0][1][0], acc.c[0][1][1], a0, b1);
        dmma884(acc.c[1][0][0], acc.c[1][0][1], a1, b0);
        dmma884(acc.c[1][1][0], acc.c[1][1][1], a1, b1);
    }
}

// visit the 8 elements of the block a lane owns: f(r, c, value&)
template <class F>
__device__ __forceinline__ void blk_foreach(BlkAcc& acc, F f, int g, int q) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) f(8 * mi + g, 8 * ni + 2 * q + e, acc.c[mi][ni][e]);
}

template <int NU, int DP>
__global__ void __launch_bounds__(FS_THREADS, 1) lml_blk_kernel(LmlSmallArgs p) {
    extern __shared__ __align__(16) double fb_smem[];
    const int N = p.N, d = p.d;
    const int nblk = (N + FB - 1) / FB, Nr = nblk * FB;
    double* Lp = fb_smem;                        // packed lower triangle: (i, j <= i) at tri(i) + j
    double* yv = Lp + size_t(N) * (N + 1) / 2;   // [N]
    double* zv = yv + N;                         // [N]
    double* al = zv + N;                         // [N]
    double* dg = al + N;                         // [N] diagonal of L (for log|L|)
    // ONE panel buffer P[Nr][FBP]: rows i below the current block hold the column panel (L[i][k-block][.]), rows j up to
    // the current block hold the row panel transposed (P[j][r] = X[k-block][r][j]) — the two never overlap
    double* P = dg + N;
    double* Dsm = P + size_t(FBP) * Nr;          // [FB][FB+1] diagonal block scratch
    double* Dinv = Dsm + FB * (FB + 1);          // [FB][FB+1] its inverse
    double* iell = Dinv + FB * (FB + 1);         // [DP]
    double* red = iell + 2 * DP + 16;            // [FS_WARPS][DP + 2]
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int b = blockIdx.x;
    const double* __restrict__ th = p.theta + int64_t(b) * (d + 2);
    const double c1 = exp(th[0]), c2 = exp(th[d + 1]);
    if (tid < DP) iell[tid] = (tid < d) ? exp(-th[1 + tid]) : 0.0;
    if (tid == 0) s_fail = 0;
    const double* __restrict__ y = p.Y + int64_t(p.obj[b]) * p.Npad;
    for (int i = tid; i < N; i += FS_THREADS) yv[i] = y[i];
    __syncthreads();
#define FB_STAMP(idx) \
    if (p.clk != nullptr && tid == 0) p.clk[int64_t(b) * 8 + (idx)] = clock64();
    FB_STAMP(0)

    // packed-triangle element access with bounds: rows / columns >= N read as 0
    auto LP = [&](int i, int j) -> double { return (i < N && j <= i) ? Lp[tri(i) + j] : 0.0; };

    // ---- 1: kernel matrix ---------------------------------------------------------------------------------------
    // 4 x 4 pair blocks on or below the diagonal, one per thread and step: per coordinate two 32-byte loads from the
    // dimension-major copy of X serve 16 pairs (consecutive threads take consecutive column blocks: coalesced)
    const int nb4 = (N + 3) >> 2, nblk4 = nb4 * (nb4 + 1) / 2;
    auto block_of = [&](int t, int& i0, int& j0) {
        int bi = int((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while (bi * (bi + 1) / 2 > t) --bi;
        while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
        i0 = 4 * bi;
        j0 = 4 * (t - bi * (bi + 1) / 2);
    };
    // squared scaled distances of the block's 16 pairs: direct differences times 1/l, as cdist(X / l, X / l) forms them
    auto block_r2 = [&](int i0, int j0, double (&r2)[4][4]) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) r2[r][c] = 0.0;
        for (int k = 0; k < d; ++k) {
            const double* row = p.XT + int64_t(k) * p.Npad;
            const double2 a01 = *reinterpret_cast<const double2*>(row + i0), a23 = *reinterpret_cast<const double2*>(row + i0 + 2);
            const double2 b01 = *reinterpret_cast<const double2*>(row + j0), b23 = *reinterpret_cast<const double2*>(row + j0 + 2);
            const double a[4] = {a01.x, a01.y, a23.x, a23.y}, bq[4] = {b01.x, b01.y, b23.x, b23.y};
            const double sc = iell[k];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double t = (a[r] - bq[c]) * sc;
                    r2[r][c] = fma(t, t, r2[r][c]);
                }
        }
    };
    for (int t = tid; t < nblk4; t += FS_THREADS) {
        int i0, j0;
        block_of(t, i0, j0);
        double r2[4][4];
        block_r2(i0, j0, r2);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = i0 + r, j = j0 + c;
                if (i < N && j <= i) {
                    double phi, gor;
                    kernel_profile<NU>(sqrt(r2[r][c]), r2[r][c], phi, gor);
                    Lp[tri(i) + j] = (i == j) ? (c1 + c2) + kJitter : fma(c1, phi, c2);
                }
            }
    }
    __syncthreads();
    FB_STAMP(1)

    // ---- 2: blocked Cholesky; the diagonal blocks end up holding L_kk^-1 ------------------------------------------
    // The sequential chain (one warp factoring the 16 x 16 diagonal block in registers, pivot column passed around by
    // shuffles) is the only thing left on the critical path: the panel below it is a triangular SOLVE with one row per
    // thread instead of "invert the block, then multiply"; with look-ahead warp 0 updates block (kb+1, kb+1) right after
    // the panel and factors it while warps 1..15 finish the trailing update; the 16 x 16 inverses that phase 3 starts
    // from are formed afterwards, all blocks at once.  zv serves as the table of 1 / L_cc until phase 4 needs it.
    double* ipv = zv;
    auto factor_diag = [&](int kb) {
        const int k0 = kb * FB;
        const long long tdiag0 = (p.clk != nullptr) ? clock64() : 0;
        const int r = lane & 15;
        double a[FB];
#pragma unroll
        for (int c = 0; c < FB; ++c) {
            double v = 0.0;
            if (c <= r) v = (k0 + r < N) ? Lp[tri(k0 + r) + k0 + c] : (r == c ? 1.0 : 0.0);  // identity padding beyond N
            a[c] = v;
        }
        bool bad = false;
#pragma unroll
        for (int c = 0; c < FB; ++c) {
            if (!bad) {
                const double dcc = __shfl_sync(0xffffffffu, a[c], c);
                if (!(dcc > 0.0)) {
                    bad = true;  // the same value in every lane: uniform
                } else {
                    const double ip = rsqrt(dcc);
                    a[c] = (r == c) ? dcc * ip : a[c] * ip;  // L_rc (rows above the pivot hold 0)
                    if (lane == c) {
                        Dinv[c * (FB + 1) + c] = ip;  // 1 / L_cc for the panel solve
                        if (k0 + c < N) ipv[k0 + c] = ip;
                    }
#pragma unroll
                    for (int cc = c + 1; cc < FB; ++cc) {
                        const double lcc = __shfl_sync(0xffffffffu, a[c], cc);  // L_{cc,c}
                        if (cc <= r) a[cc] = fma(-a[c], lcc, a[cc]);
                    }
                }
            }
        }
        if (bad) {
            if (lane == 0) s_fail = 1;
        } else if (lane < FB) {
#pragma unroll
            for (int c = 0; c < FB; ++c) {
                Dsm[lane * (FB + 1) + c] = a[c];
                if (c <= lane && k0 + lane < N) Lp[tri(k0 + lane) + k0 + c] = a[c];
            }
            if (k0 + lane < N) dg[k0 + lane] = a[lane];
        }
        if (p.clk != nullptr && lane == 0) p.clk[int64_t(b) * 8 + 6] = (kb == 0 ? 0 : p.clk[int64_t(b) * 8 + 6]) + (clock64() - tdiag0);
    };
    if (warp == 0) factor_diag(0);
    __syncthreads();
    for (int kb = 0; kb < nblk && !s_fail; ++kb) {
        const int k0 = kb * FB;
        // panel: L[i][kb-block] = A[i][kb-block] L_kk^-T by substitution, one row per thread; also into the panel buffer
        {
            const int i = k0 + FB + tid;
            if (i < Nr) {
                double x[FB];
#pragma unroll
                for (int c = 0; c < FB; ++c) {
                    double sacc = LP(i, k0 + c);
#pragma unroll
                    for (int j = 0; j < c; ++j) sacc = fma(-x[j], Dsm[c * (FB + 1) + j], sacc);
                    x[c] = sacc * Dinv[c * (FB + 1) + c];
                }
#pragma unroll
                for (int c = 0; c < FB; ++c) {
                    const bool in = i < N && k0 + c < N;
                    if (in) Lp[tri(i) + k0 + c] = x[c];
                    P[i * FBP + c] = in ? x[c] : 0.0;
                }
            }
        }
        __syncthreads();
        // trailing update: A[ib][jb] -= L[ib][kb] L[jb][kb]^T for kb < jb <= ib; pair 0 is the next diagonal block
        {
            const int nt = nblk - kb - 1;
            const int npair = nt * (nt + 1) / 2;
            auto do_pair = [&](int pr) {
                int bi = int((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
                while (bi * (bi + 1) / 2 > pr) --bi;
                while ((bi + 1) * (bi + 2) / 2 <= pr) ++bi;
                const int bj = pr - bi * (bi + 1) / 2;
                const int i0 = (kb + 1 + bi) * FB, j0 = (kb + 1 + bj) * FB;
                BlkAcc acc;
                acc.zero();
                blk_mma(acc, [&](int r, int k) { return P[(i0 + r) * FBP + k]; },
                        [&](int k, int n) { return P[(j0 + n) * FBP + k]; }, g, q);
                blk_foreach(acc, [&](int r, int c, double& v) {
                    const int i = i0 + r, j = j0 + c;
                    if (i < N && j <= i) Lp[tri(i) + j] -= v;
                }, g, q);
            };
            if (warp == 0) {
                if (npair > 0) {
                    do_pair(0);
                    __syncwarp();
                    factor_diag(kb + 1);  // overwrites Dsm / Dinv: the panel solve of step kb is behind the barrier above
                }
            } else {
                for (int pr = warp; pr < npair; pr += FS_WARPS - 1) do_pair(pr);
            }
        }
        __syncthreads();
    }
    // inverses of all diagonal blocks, one warp per block, one column per lane (forward substitution in registers)
    if (!s_fail) {
        const long long tinv0 = (p.clk != nullptr) ? clock64() : 0;
        for (int kb = warp; kb < nblk; kb += FS_WARPS) {
            const int k0 = kb * FB;
            double x[FB];
            const int j = lane & 15;
            auto Lkk = [&](int i, int k) -> double {  // the factored block with its identity padding
                return (k0 + i < N) ? Lp[tri(k0 + i) + k0 + k] : (i == k ? 1.0 : 0.0);
            };
#pragma unroll
            for (int i = 0; i < FB; ++i) x[i] = 0.0;
#pragma unroll
            for (int i = 0; i < FB; ++i) {
                if (i < j) continue;
                const double idiag = (k0 + i < N) ? ipv[k0 + i] : 1.0;
                if (i == j) {
                    x[i] = idiag;
                } else {
                    double sacc = 0.0;
#pragma unroll
                    for (int k = 0; k < FB; ++k)
                        if (k >= j && k < i) sacc = fma(Lkk(i, k), x[k], sacc);
                    x[i] = -sacc * idiag;
                }
            }
            __syncwarp();  // every lane has read the block before anybody overwrites it
            if (lane < FB) {
#pragma unroll
                for (int i = 0; i < FB; ++i)
                    if (i >= j && k0 + i < N) Lp[tri(k0 + i) + k0 + j] = x[i];
            }
        }
        if (p.clk != nullptr && tid == 0) p.clk[int64_t(b) * 8 + 7] = clock64() - tinv0;
    }
    __syncthreads();
    double* out = p.res + int64_t(b) * (d + 3);
    double logdet = 0.0;
    if (warp == 0 && !s_fail) {
        for (int k = lane; k < N; k += 32) logdet += log(dg[k]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) logdet += __shfl_xor_sync(0xffffffffu, logdet, off);
    }
    if (s_fail) {
        if (tid == 0) out[0] = -INFINITY;
        if (tid < d + 2) out[1 + tid] = 0.0;
        return;
    }
    FB_STAMP(2)

    // ---- 3: X = L^-1 in place (block rows top to bottom) -------------------------------------------------------------
    // invariant before step kb: block rows < kb hold X; rows >= kb hold R = I - (partial products) in columns < kb,
    // L in columns >= kb (strictly lower blocks) and L_kk^-1 on the diagonal
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * FB;
        // (a) row block kb: X[kb][jb] = Dinv_kb * R[kb][jb] (jb < kb); X[kb][kb] = Dinv_kb.  Result -> rowP (and storage)
        for (int jb = warp; jb <= kb; jb += FS_WARPS) {
            const int j0 = jb * FB;
            BlkAcc acc;
            acc.zero();
            if (jb < kb) {
                blk_mma(acc, [&](int r, int k) { return (k <= r) ? LP(k0 + r, k0 + k) : 0.0; },
                        [&](int k, int n) { return LP(k0 + k, j0 + n); }, g, q);
            } else {
                blk_foreach(acc, [&](int r, int c, double& v) { v = (c <= r) ? LP(k0 + r, k0 + c) : 0.0; }, g, q);
            }
            __syncwarp();
            blk_foreach(acc, [&](int r, int c, double& v) {
                if (jb < kb && k0 + r < N) Lp[tri(k0 + r) + j0 + c] = v;
                P[(j0 + c) * FBP + r] = (k0 + r < N) ? v : 0.0;
            }, g, q);
        }
        // (b) column panel L[ib][kb] for ib > kb -> colP
        for (int e = tid; e < (Nr - k0 - FB) * FB; e += FS_THREADS) {
            const int i = k0 + FB + e / FB, c = e % FB;
            P[i * FBP + c] = LP(i, k0 + c);
        }
        __syncthreads();
        // (c) R[ib][jb] = (jb == kb ? 0 : R[ib][jb]) - L[ib][kb] * X[kb][jb]   for ib > kb, jb <= kb
        {
            const int nrow = nblk - kb - 1, ncol = kb + 1;
            for (int pr = warp; pr < nrow * ncol; pr += FS_WARPS) {
                const int ib = kb + 1 + pr / ncol, jb = pr % ncol;
                const int i0 = ib * FB, j0 = jb * FB;
                BlkAcc acc;
                acc.zero();
                blk_mma(acc, [&](int r, int k) { return P[(i0 + r) * FBP + k]; },
                        [&](int k, int n) { return P[(j0 + n) * FBP + k]; }, g, q);
                blk_foreach(acc, [&](int r, int c, double& v) {
                    const int i = i0 + r, j = j0 + c;
                    if (i < N && j < N) Lp[tri(i) + j] = (jb == kb ? 0.0 : Lp[tri(i) + j]) - v;
                }, g, q);
            }
        }
        __syncthreads();
    }
    FB_STAMP(3)

    // ---- 4: z = X y, alpha = X^T z ------------------------------------------------------------------------------------
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
        double qq = 0.0;
        for (int i = lane; i < N; i += 32) qq = fma(zv[i], zv[i], qq);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, off);
        if (lane == 0) out[0] = -0.5 * qq - logdet - 0.5 * N * 1.8378770664093453;
    }
    __syncthreads();
    FB_STAMP(4)
    if (!p.want_grad) return;

    // ---- 5a: G = X^T X in place (block rows top to bottom) ---------------------------------------------------------------
    for (int tb = 0; tb < nblk; ++tb) {
        const int t0 = tb * FB;
        // row panel X[tb][0 .. tb] -> rowP (the diagonal block is lower triangular)
        for (int e = tid; e < FB * (t0 + FB); e += FS_THREADS) {
            const int r = e / (t0 + FB), j = e % (t0 + FB);
            P[j * FBP + r] = LP(t0 + r, j);
        }
        __syncthreads();
        {
            const int nb = tb + 1, npair = nb * (nb + 1) / 2;
            for (int pr = warp; pr < npair; pr += FS_WARPS) {
                int ib = int((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
                while (ib * (ib + 1) / 2 > pr) --ib;
                while ((ib + 1) * (ib + 2) / 2 <= pr) ++ib;
                const int jb = pr - ib * (ib + 1) / 2;
                const int i0 = ib * FB, j0 = jb * FB;
                BlkAcc acc;
                acc.zero();
                blk_mma(acc, [&](int r, int k) { return P[(i0 + r) * FBP + k]; },
                        [&](int k, int n) { return P[(j0 + n) * FBP + k]; }, g, q);
                blk_foreach(acc, [&](int r, int c, double& v) {
                    const int i = i0 + r, j = j0 + c;
                    if (i < N && j <= i) Lp[tri(i) + j] = (ib == tb ? 0.0 : Lp[tri(i) + j]) + v;
                }, g, q);
            }
        }
        __syncthreads();
    }

    // ---- 5b: gradient contraction ---------------------------------------------------------------------------------------
    // the same 4 x 4 pair blocks: distances, then the weights w (alpha alpha^T - K^-1, off-diagonal pairs twice) times the
    // kernel derivatives, then a second sweep over the coordinates for the length-scale components
    double gacc[DP + 2];
#pragma unroll
    for (int k = 0; k < DP + 2; ++k) gacc[k] = 0.0;
    for (int t = tid; t < nblk4; t += FS_THREADS) {
        int i0, j0;
        block_of(t, i0, j0);
        double wh[4][4];  // squared distance first, then w * c1 * h(r)
        block_r2(i0, j0, wh);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = i0 + r, j = j0 + c;
                double whv = 0.0;
                if (i < N && j <= i) {
                    double w = al[i] * al[j] - Lp[tri(i) + j];
                    if (i != j) w *= 2.0;
                    const double r2v = wh[r][c], rr = sqrt(r2v);
                    double phi, gor;
                    kernel_profile<NU>(rr, r2v, phi, gor);
                    if (i == j) phi = 1.0;
                    gacc[0] = fma(w * c1, phi, gacc[0]);
                    gacc[DP + 1] = fma(w, c2, gacc[DP + 1]);
                    whv = w * c1 * h_profile<NU>(rr, r2v, phi);
                }
                wh[r][c] = whv;
            }
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            if (k < d) {
                const double* row = p.XT + int64_t(k) * p.Npad;
                const double2 a01 = *reinterpret_cast<const double2*>(row + i0), a23 = *reinterpret_cast<const double2*>(row + i0 + 2);
                const double2 b01 = *reinterpret_cast<const double2*>(row + j0), b23 = *reinterpret_cast<const double2*>(row + j0 + 2);
                const double a[4] = {a01.x, a01.y, a23.x, a23.y}, bq[4] = {b01.x, b01.y, b23.x, b23.y};
                const double sc = iell[k];
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; c += 2) {
                        const double t0 = (a[r] - bq[c]) * sc, t1 = (a[r] - bq[c + 1]) * sc;
                        s0 = fma(wh[r][c], t0 * t0, s0);
                        s1 = fma(wh[r][c + 1], t1 * t1, s1);
                    }
                gacc[1 + k] += s0 + s1;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < DP + 2; ++k) {
        double v = gacc[k];
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
    FB_STAMP(5)
#undef FB_STAMP
}

}  // namespace aoed
