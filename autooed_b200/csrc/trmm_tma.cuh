// Persistent, warp-specialised FP64 triangular product for sm_100a: TMA + mbarrier ring + DMMA.
//
// Same contract as trmm_kernel (trmm.cuh):  Out[b][i][c] = sum_k A[b][i][k] * B[b][k][c] over the triangular k-range
// of row tile i, optional per-column sums of squares.  What changed is the machinery around the DMMA loop:
//
//   * one CTA per SM, alive for the whole launch; tiles (batch, column tile, row tile) are handed out by an atomic
//     counter in heaviest-first order (greedy longest-processing-time), so there is no per-CTA prologue bubble and no tail;
//   * ONE producer lane (in a warpgroup of its own that hands its registers to the consumers with setmaxnreg) issues cp.async.bulk.tensor (TMA) loads of the A tile [128 x 16] and the B tile
//     [16 x 128] (as eight [16 x 16] boxes) into a 6-stage shared-memory ring, completion on `full` mbarriers;
//   * 16 consumer warps (4 x 4 sub-tiles of 32 x 32) wait on `full`, issue their 64 DMMA.8x8x4 and arrive on `empty` —
//     there is NO CTA-wide barrier in the main loop, warps drift freely (a warp that skips k-steps above the diagonal
//     runs ahead instead of idling at a __syncthreads), and the producer keeps prefetching across tile boundaries while
//     the consumers store their accumulators.
//
// Shared-memory layout is whatever TMA writes with CU_TENSOR_MAP_SWIZZLE_128B (16-byte chunk index XOR (row & 7) inside
// each 1024-byte group); the fragment -> (row, k, column) assignment of the m8n8k4 MMA is permuted so that every fragment
// load is a conflict-free LDS.128:
//     k of MMA step s for lane quad-index q :  kappa(q, s) = 2q + 8 (s >> 1) + (s & 1)
//     row of lane group g inside an 8-row block : rho(g) = ((g & 1) << 2) | (g >> 1)
//     column j of 8-column block nb : 2 j + (nb & 1) + 16 (nb >> 1)     (=> 4 contiguous output columns per lane)
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "trmm.cuh"

namespace aoed {

constexpr int PT_STAGES = 6;
constexpr int PT_CONSUMER_WARPS = 16;
constexpr int PT_THREADS = (PT_CONSUMER_WARPS + 4) * 32;  // + one producer warpgroup (setmaxnreg works per 4 warps)
constexpr int PT_REGS_PRODUCER = 24, PT_REGS_CONSUMER = 112;  // pool = regs released by the producer WG: 640*96 >= 512*112 + 128*24
constexpr int PT_A_BYTES = TM * TK * 8;   // 16 KB, rows of 128 B
constexpr int PT_B_BYTES = TK * 128 * 8;  // 16 KB, eight [16 k][16 col] boxes of 2 KB
constexpr int PT_STAGE_BYTES = PT_A_BYTES + PT_B_BYTES;
constexpr size_t PT_SMEM = size_t(PT_STAGES) * PT_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers, tile ids*/;
constexpr int PT_SS_PER_TILE = 4;  // sum-of-squares partials per row tile: one per 32-row warp slice

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Blocks until the phase with the given parity has completed.  Bounded: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* dbg = nullptr, int where = 0, int info = 0) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x3E8;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) {  // tens of seconds of 1 us suspends: only a protocol bug gets here
            if (dbg != nullptr && atomicCAS(dbg, 0, where) == 0) {
                dbg[1] = blockIdx.x; dbg[2] = threadIdx.x; dbg[3] = info; dbg[4] = int(parity);
                __threadfence_system();
            }
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

struct TrmmTmaArgs {
    double* Out;      // [batch][mTiles*TM][ldo] or nullptr
    int64_t strideOut;
    int ldo;
    double* sumsq;    // [batch][mTiles*4][ldo] or nullptr
    int64_t strideSS;
    int M;            // valid rows
    int Kpad;         // valid k rounded up to TK
    int Npad;         // rows per batch slab in both tensor maps
    int mTiles, colTiles, batch;
    int* counter;     // tile ticket counter, zeroed before the launch
    int* dbg;         // host-mapped diagnostics (who timed out where), or nullptr
    int colGroup;     // column tiles whose B panels (all batches) are meant to stay L2-resident together
    // fused gradient moments of the variance (UPPER product only; replaces gradsum_kernel, gp.py:201-205): with
    // p = Gr[row][col] * Out[row][col], per 32-row warp slice and column  sum_rows p  and  sum_rows p * Ts[row][k], k < dims
    const double* Gr;   // [batch][rows][ldo] (same strides as Out) or nullptr
    const double* Ts;   // [batch][Npad][DPs] training points divided by the length scales
    double* part2;      // [batch][mTiles * 4][1 + DPs][ldo]
    int DPs, dims;
};

constexpr int PT_COL_GROUP = 8;  // default colGroup (AOED_TRMM_COLGROUP overrides)

// Ticket -> tile.  Order: groups of PT_COL_GROUP column tiles outermost (their B panels, 8*128 columns x Npad x batch
// = 50 MB at N = 2048, m = 3, stay in the 126 MB L2 while every row tile of the group streams over them), inside a group
// heaviest row tiles first, then batch, then column tile (CTAs running concurrently share the A panel of one row tile).
template <bool UPPER>
__device__ __forceinline__ void decode_tile(int t, int mTiles, int colTiles, int batch, int colGroup, int& mt, int& b, int& ct) {
    const int full = colGroup * batch * mTiles;
    const int ng = (colTiles + colGroup - 1) / colGroup;
    int grp = t / full, r = t - grp * full, gsz = colGroup;
    if (grp >= ng - 1) {
        grp = ng - 1;
        r = t - grp * full;
        gsz = colTiles - grp * colGroup;
    }
    const int mr = r / (gsz * batch), rem = r - mr * (gsz * batch);
    mt = UPPER ? mr : mTiles - 1 - mr;
    b = rem / gsz;
    ct = grp * colGroup + (rem - b * gsz);
}

template <bool UPPER>
__global__ void __launch_bounds__(PT_THREADS, 1)
trmm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TrmmTmaArgs p) {
    extern __shared__ uint8_t pt_raw[];
    const uint32_t base = (smem_u32(pt_raw) + 1023u) & ~1023u;  // SWIZZLE_128B pattern follows absolute address bits
    uint8_t* gen = pt_raw + (base - smem_u32(pt_raw));
    const uint32_t bars = base + PT_STAGES * PT_STAGE_BYTES;  // full[0..S), empty[0..S)
    volatile int* stage_tile = reinterpret_cast<volatile int*>(gen + PT_STAGES * PT_STAGE_BYTES + 2 * PT_STAGES * 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = p.mTiles * p.colTiles * p.batch;

    if (tid == 0) {
        for (int s = 0; s < PT_STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);                                   // full: producer's arrive + TMA bytes
            mbar_init(bars + 8 * (PT_STAGES + s), PT_CONSUMER_WARPS);     // empty: one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= PT_CONSUMER_WARPS) {
        // ================================ producer ================================
        // give the registers of the (almost idle) producer warpgroup to the 16 consumer warps
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PT_REGS_PRODUCER));
        if (warp != PT_CONSUMER_WARPS || lane != 0) return;
        uint32_t it = 0;
        for (;;) {
            const int tile = atomicAdd(p.counter, 1);
            if (tile >= total) {
                const uint32_t s = it % PT_STAGES, r = it / PT_STAGES;
                mbar_wait(bars + 8 * (PT_STAGES + s), (r & 1) ^ 1, p.dbg, 1, int(it));
                stage_tile[s] = -1;
                mbar_arrive(bars + 8 * s);
                return;
            }
            int mt, b, ct;
            decode_tile<UPPER>(tile, p.mTiles, p.colTiles, p.batch, p.colGroup, mt, b, ct);
            const int k_begin = UPPER ? mt * TM : 0;
            const int k_end = UPPER ? p.Kpad : min(p.Kpad, (mt + 1) * TM);
            const int KT = (k_end - k_begin) / TK;
            const int rowA = b * p.Npad + mt * TM, rowB0 = b * p.Npad + k_begin, colB = ct * 128;
            for (int kt = 0; kt < KT; ++kt, ++it) {
                const uint32_t s = it % PT_STAGES, r = it / PT_STAGES;
                mbar_wait(bars + 8 * (PT_STAGES + s), (r & 1) ^ 1, p.dbg, 2, int(it));
                stage_tile[s] = tile;
                const uint32_t full = bars + 8 * s, dstA = base + s * PT_STAGE_BYTES, dstB = dstA + PT_A_BYTES;
                mbar_arrive_expect_tx(full, PT_STAGE_BYTES);
                tma_load_2d(dstA, &mapA, full, k_begin + kt * TK, rowA);
#pragma unroll
                for (int cg = 0; cg < 8; ++cg) tma_load_2d(dstB + cg * 2048, &mapB, full, colB + cg * 16, rowB0 + kt * TK);
            }
        }
    }

    // ================================ consumers ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PT_REGS_CONSUMER));
    const int wm = warp >> 2, wn = warp & 3;
    const int g = lane >> 2, q = lane & 3;
    const int rho = ((g & 1) << 2) | (g >> 1);
    // byte offsets inside a stage (swizzled): A row (wm*32 + mb*8 + rho), chunks q and q+4
    const uint32_t offA0 = uint32_t(wm * 32 + rho) * 128u + (uint32_t(q ^ rho) << 4);
    const uint32_t offA1 = uint32_t(wm * 32 + rho) * 128u + (uint32_t((q + 4) ^ rho) << 4);
    // B: box (wn*2 + pp), row kappa(q, s), chunk g ^ (kappa & 7)
    uint32_t offB[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int k = 2 * q + 8 * (s >> 1) + (s & 1);
        offB[s] = PT_A_BYTES + uint32_t(wn * 2) * 2048u + uint32_t(k) * 128u + (uint32_t(g ^ (k & 7)) << 4);
    }

    uint32_t it = 0;
    for (;;) {
        {
            const uint32_t s = it % PT_STAGES, r = it / PT_STAGES;
            mbar_wait(bars + 8 * s, r & 1, p.dbg, 3, int(it));
        }
        const int tile = stage_tile[it % PT_STAGES];
        if (tile < 0) break;
        int mt, b, ct;
        decode_tile<UPPER>(tile, p.mTiles, p.colTiles, p.batch, p.colGroup, mt, b, ct);
        const int k_begin = UPPER ? mt * TM : 0;
        const int k_end = UPPER ? p.Kpad : min(p.Kpad, (mt + 1) * TM);
        const int KT = (k_end - k_begin) / TK;
        const int diag0 = UPPER ? 0 : mt * (TM / TK);
        const bool rows_live = mt * TM + wm * 32 < p.M;

        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kt = 0; kt < KT; ++kt, ++it) {
            const uint32_t s = it % PT_STAGES, r = it / PT_STAGES;
            if (kt > 0) mbar_wait(bars + 8 * s, r & 1, p.dbg, 4, int(it));
            // triangular skip inside the diagonal 128-block, and row slices that are pure padding
            const int kd = kt - diag0;
            bool active = rows_live;
            if (UPPER) {
                if (kd < 2 * wm) active = false;
            } else {
                if (kd > 2 * wm + 1) active = false;
            }
            if (active) {
                const uint8_t* st = gen + s * PT_STAGE_BYTES;
                double2 a2[4][2], b2[2][4];
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) {
                    a2[mb][0] = *reinterpret_cast<const double2*>(st + offA0 + mb * 1024);
                    a2[mb][1] = *reinterpret_cast<const double2*>(st + offA1 + mb * 1024);
                }
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

        // ---- epilogue: lane owns, per (mb, pp), 4 contiguous columns  col0 = wn*32 + 16 pp + 4 q ----
        if (p.Out != nullptr) {
            double* __restrict__ Out = p.Out + b * p.strideOut;
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                const int row = mt * TM + wm * 32 + mb * 8 + rho;
                if (row < p.M) {
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        double* dst = Out + int64_t(row) * p.ldo + ct * 128 + wn * 32 + 16 * pp + 4 * q;
                        *reinterpret_cast<double2*>(dst) = make_double2(acc[mb][2 * pp][0], acc[mb][2 * pp + 1][0]);
                        *reinterpret_cast<double2*>(dst + 2) = make_double2(acc[mb][2 * pp][1], acc[mb][2 * pp + 1][1]);
                    }
                }
            }
        }
        if (UPPER && p.part2 != nullptr) {
            // ---- fused variance-gradient moments: the warp's 32 x 32 sub-tile of V never has to be re-read from HBM ----
            const double* __restrict__ Gr = p.Gr + b * p.strideOut;
            const double* __restrict__ Ts = p.Ts + int64_t(b) * p.Npad * p.DPs;
            const int row0 = mt * TM + wm * 32 + rho;
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {  // acc <- Gr * V (rows in the zero padding contribute nothing)
                const int row = row0 + mb * 8;
                const bool live = row < p.M;
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    double2 g0 = make_double2(0.0, 0.0), g1 = g0;
                    if (live) {
                        const double* src = Gr + int64_t(row) * p.ldo + ct * 128 + wn * 32 + 16 * pp + 4 * q;
                        g0 = *reinterpret_cast<const double2*>(src);
                        g1 = *reinterpret_cast<const double2*>(src + 2);
                    }
                    acc[mb][2 * pp][0] *= g0.x;
                    acc[mb][2 * pp + 1][0] *= g0.y;
                    acc[mb][2 * pp][1] *= g1.x;
                    acc[mb][2 * pp + 1][1] *= g1.y;
                }
            }
            double* __restrict__ P2 = p.part2 + (int64_t(b) * p.mTiles * PT_SS_PER_TILE + mt * PT_SS_PER_TILE + wm) * int64_t(1 + p.DPs) * p.ldo +
                                      ct * 128 + wn * 32 + 16 * (g >> 2) + 4 * q + (g & 3);
            // three moments per pass: their twelve Ts loads (L2 hits, ~500 cycles) are in flight together
            for (int kk0 = 0; kk0 <= p.dims; kk0 += 3) {
                double t[3][4];
#pragma unroll
                for (int u = 0; u < 3; ++u)
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) {
                        const int row = row0 + mb * 8, kk = kk0 + u;
                        t[u][mb] = (kk == 0) ? 1.0 : ((row < p.M && kk <= p.dims) ? Ts[int64_t(row) * p.DPs + kk - 1] : 0.0);
                    }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int kk = kk0 + u;
                    // s8[j]: column j = 4 pp + 2 e + (nb & 1) of the lane's eight, summed over its four rows
                    double s8[8];
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const int nb = 2 * pp + (j4 & 1), e = j4 >> 1;
                            double v = acc[0][nb][e] * t[u][0];
#pragma unroll
                            for (int mb = 1; mb < 4; ++mb) v = fma(acc[mb][nb][e], t[u][mb], v);
                            s8[4 * pp + j4] = v;
                        }
                    // reduce over the eight lane groups (rows) and scatter: lane group g ends up with column j = g
                    double s4[4], s2[2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double give = (g & 4) ? s8[j] : s8[j + 4];
                        const double keep = (g & 4) ? s8[j + 4] : s8[j];
                        s4[j] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const double give = (g & 2) ? s4[j] : s4[j + 2];
                        const double keep = (g & 2) ? s4[j + 2] : s4[j];
                        s2[j] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
                    }
                    const double give = (g & 1) ? s2[0] : s2[1];
                    const double keep = (g & 1) ? s2[1] : s2[0];
                    const double tot = keep + __shfl_xor_sync(0xffffffffu, give, 4);
                    if (kk <= p.dims) P2[int64_t(kk) * p.ldo] = tot;
                }
            }
        }
        if (p.sumsq != nullptr) {
            double* __restrict__ SS = p.sumsq + b * p.strideSS + int64_t(mt * PT_SS_PER_TILE + wm) * p.ldo + ct * 128 + wn * 32;
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                double ss[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int nb = 2 * pp + (j & 1), e = j >> 1;  // column offset j = 2 e + (nb & 1)
                    double v = 0.0;
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) v = fma(acc[mb][nb][e], acc[mb][nb][e], v);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    ss[j] = v;
                }
                if (g == 0) {
                    double* dst = SS + 16 * pp + 4 * q;
                    *reinterpret_cast<double2*>(dst) = make_double2(ss[0], ss[1]);
                    *reinterpret_cast<double2*>(dst + 2) = make_double2(ss[2], ss[3]);
                }
            }
        }
    }
}

}  // namespace aoed
