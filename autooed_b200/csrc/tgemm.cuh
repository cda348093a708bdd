// Table-driven persistent FP64 tile GEMM for sm_100a (TMA + mbarrier ring + DMMA): the building block of the blocked
// Cholesky factorisation, the triangular inverse and K^-1 = L^-T L^-1 of the hyper-parameter fit (blocked.cu).
//
// One launch processes `ntiles` 128 x 128 output tiles of EACH of `batch` independent problems.  A tile is described
// by a TileDesc (built once per matrix size on the host): where its A and B operand strips start, how many 16-wide
// k-steps it sums over, where the result goes and how it is combined.  Ticket t -> (tile t / batch, problem t % batch);
// the table is sorted heaviest first, so the atomic ticket counter gives a longest-processing-time schedule.
//
//   NT = false:  Out[r][c] (+)= sum_k A[r][k] * B[k][c]     A row-major (k contiguous), B k-major   (as trmm_tma_kernel)
//   NT = true :  Out[r][c] (+)= sum_k A[r][k] * B[c][k]     both row-major with k contiguous        (L L^T products)
//
// Machinery and fragment layout are those of trmm_tma_kernel (trmm_tma.cuh): one producer lane issuing
// cp.async.bulk.tensor into a 6-stage ring, 16 consumer warps with 32 x 32 sub-tiles of m8n8k4 DMMAs, no CTA-wide
// barrier in the main loop.  In NT mode the B strip is loaded with the A-style box (128 rows x 16 k) and read with
// the A-style conflict-free LDS.128 pattern.
#pragma once
#include "trmm_tma.cuh"

namespace aoed {

enum : int { TG_WRITE = 0, TG_NEG = 1, TG_SUB = 2 };  // Out = acc | Out = -acc | Out = Out - acc

struct TileDesc {
    int rowA, colA;    // TMA coordinates (row inside one problem, first k column) of the A strip
    int rowB, colB;    // NN: first k row and column of the B strip; NT: row (= output column index) and first k column
    int KT;            // number of 16-wide k-steps (>= 1)
    int mode;          // TG_*
    int outRow, outCol;  // position of the output tile inside one problem's matrix
};

struct TgemmArgs {
    const TileDesc* tiles;
    int ntiles, batch;
    int rowsPerProblem;  // row offset between consecutive problems in both tensor maps (Npad)
    double* Out;
    int64_t strideOut;
    int ldo;
    int* counter;  // ticket counter, zero before the launch
    int* dbg;
};

template <bool NT>
__global__ void __launch_bounds__(PT_THREADS, 1)
tgemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TgemmArgs p) {
    extern __shared__ uint8_t pt_raw[];
    const uint32_t base = (smem_u32(pt_raw) + 1023u) & ~1023u;
    uint8_t* gen = pt_raw + (base - smem_u32(pt_raw));
    const uint32_t bars = base + PT_STAGES * PT_STAGE_BYTES;
    volatile int* stage_tile = reinterpret_cast<volatile int*>(gen + PT_STAGES * PT_STAGE_BYTES + 2 * PT_STAGES * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = p.ntiles * p.batch;

    if (tid == 0) {
        for (int s = 0; s < PT_STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (PT_STAGES + s), PT_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= PT_CONSUMER_WARPS) {
        // ================================ producer ================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PT_REGS_PRODUCER));
        if (warp != PT_CONSUMER_WARPS || lane != 0) return;
        uint32_t it = 0;
        for (;;) {
            const int ticket = atomicAdd(p.counter, 1);
            if (ticket >= total) {
                const uint32_t s = it % PT_STAGES, r = it / PT_STAGES;
                mbar_wait(bars + 8 * (PT_STAGES + s), (r & 1) ^ 1, p.dbg, 1, int(it));
                stage_tile[s] = -1;
                mbar_arrive(bars + 8 * s);
                return;
            }
            const int ti = ticket / p.batch, b = ticket - ti * p.batch;
            const TileDesc d = p.tiles[ti];
            const int rowA = b * p.rowsPerProblem + d.rowA, rowB = b * p.rowsPerProblem + d.rowB;
            for (int kt = 0; kt < d.KT; ++kt, ++it) {
                const uint32_t s = it % PT_STAGES, r = it / PT_STAGES;
                mbar_wait(bars + 8 * (PT_STAGES + s), (r & 1) ^ 1, p.dbg, 2, int(it));
                stage_tile[s] = ticket;
                const uint32_t full = bars + 8 * s, dstA = base + s * PT_STAGE_BYTES, dstB = dstA + PT_A_BYTES;
                mbar_arrive_expect_tx(full, PT_STAGE_BYTES);
                tma_load_2d(dstA, &mapA, full, d.colA + kt * TK, rowA);
                if (NT) {
                    tma_load_2d(dstB, &mapB, full, d.colB + kt * TK, rowB);
                } else {
#pragma unroll
                    for (int cg = 0; cg < 8; ++cg) tma_load_2d(dstB + cg * 2048, &mapB, full, d.colB + cg * 16, rowB + kt * TK);
                }
            }
        }
    }

    // ================================ consumers ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PT_REGS_CONSUMER));
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, q = lane & 3;
    const int rho = ((g & 1) << 2) | (g >> 1);
    const uint32_t offA0 = uint32_t(wm * 32 + rho) * 128u + (uint32_t(q ^ rho) << 4);
    const uint32_t offA1 = uint32_t(wm * 32 + rho) * 128u + (uint32_t((q + 4) ^ rho) << 4);
    // NN: B box (wn*2 + pp), row kappa(q, s), chunk g ^ (kappa & 7);  NT: B row (wn*32 + nb*8 + rho), chunks q and q + 4
    uint32_t offB[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int k = 2 * q + 8 * (s >> 1) + (s & 1);
        offB[s] = PT_A_BYTES + uint32_t(wn * 2) * 2048u + uint32_t(k) * 128u + (uint32_t(g ^ (k & 7)) << 4);
    }
    const uint32_t offBn0 = PT_A_BYTES + uint32_t(wn * 32 + rho) * 128u + (uint32_t(q ^ rho) << 4);
    const uint32_t offBn1 = PT_A_BYTES + uint32_t(wn * 32 + rho) * 128u + (uint32_t((q + 4) ^ rho) << 4);

    uint32_t it = 0;
    for (;;) {
        {
            const uint32_t s = it % PT_STAGES, r = it / PT_STAGES;
            mbar_wait(bars + 8 * s, r & 1, p.dbg, 3, int(it));
        }
        const int ticket = stage_tile[it % PT_STAGES];
        if (ticket < 0) break;
        const int ti = ticket / p.batch, b = ticket - ti * p.batch;
        const TileDesc d = p.tiles[ti];

        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = 0; kt < d.KT; ++kt, ++it) {
            const uint32_t s = it % PT_STAGES, r = it / PT_STAGES;
            if (kt > 0) mbar_wait(bars + 8 * s, r & 1, p.dbg, 4, int(it));
            const uint8_t* st = gen + s * PT_STAGE_BYTES;
            double2 a2[4][2];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                a2[mb][0] = *reinterpret_cast<const double2*>(st + offA0 + mb * 1024);
                a2[mb][1] = *reinterpret_cast<const double2*>(st + offA1 + mb * 1024);
            }
            if (NT) {
                double2 bn[4][2];
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    bn[nb][0] = *reinterpret_cast<const double2*>(st + offBn0 + nb * 1024);
                    bn[nb][1] = *reinterpret_cast<const double2*>(st + offBn1 + nb * 1024);
                }
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4)
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) {
                        const double av = (s4 & 1) ? a2[mb][s4 >> 1].y : a2[mb][s4 >> 1].x;
#pragma unroll
                        for (int nb = 0; nb < 4; ++nb) {
                            const double bv = (s4 & 1) ? bn[nb][s4 >> 1].y : bn[nb][s4 >> 1].x;
                            dmma884(acc[mb][nb][0], acc[mb][nb][1], av, bv);
                        }
                    }
            } else {
                double2 b2[2][4];
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    b2[0][s4] = *reinterpret_cast<const double2*>(st + offB[s4]);
                    b2[1][s4] = *reinterpret_cast<const double2*>(st + offB[s4] + 2048);
                }
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4)
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) {
                        const double av = (s4 & 1) ? a2[mb][s4 >> 1].y : a2[mb][s4 >> 1].x;
#pragma unroll
                        for (int nb = 0; nb < 4; ++nb) {
                            const double bv = (nb & 1) ? b2[nb >> 1][s4].y : b2[nb >> 1][s4].x;
                            dmma884(acc[mb][nb][0], acc[mb][nb][1], av, bv);
                        }
                    }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (PT_STAGES + s));
        }

        // ---- epilogue ----
        double* __restrict__ Out = p.Out + b * p.strideOut + int64_t(d.outRow) * p.ldo + d.outCol;
        if (NT) {
            // lane owns rows wm*32 + mb*8 + rho, columns wn*32 + nb*8 + q (e = 0) and + 4 + q (e = 1)
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                double* rowp = Out + int64_t(wm * 32 + mb * 8 + rho) * p.ldo + wn * 32 + q;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double* dst = rowp + nb * 8 + 4 * e;
                        const double v = acc[mb][nb][e];
                        if (d.mode == TG_SUB) *dst = *dst - v;
                        else *dst = (d.mode == TG_NEG) ? -v : v;
                    }
            }
        } else {
            // lane owns, per (mb, pp), 4 contiguous columns  wn*32 + 16 pp + 4 q
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                double* rowp = Out + int64_t(wm * 32 + mb * 8 + rho) * p.ldo + wn * 32 + 4 * q;
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    double2* dst = reinterpret_cast<double2*>(rowp + 16 * pp);
                    double2 v0 = make_double2(acc[mb][2 * pp][0], acc[mb][2 * pp + 1][0]);
                    double2 v1 = make_double2(acc[mb][2 * pp][1], acc[mb][2 * pp + 1][1]);
                    if (d.mode == TG_SUB) {
                        const double2 c0 = dst[0], c1 = dst[1];
                        v0 = make_double2(c0.x - v0.x, c0.y - v0.y);
                        v1 = make_double2(c1.x - v1.x, c1.y - v1.y);
                    } else if (d.mode == TG_NEG) {
                        v0 = make_double2(-v0.x, -v0.y);
                        v1 = make_double2(-v1.x, -v1.y);
                    }
                    dst[0] = v0;
                    dst[1] = v1;
                }
            }
        }
    }
}

}  // namespace aoed
