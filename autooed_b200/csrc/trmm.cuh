// Batched FP64 triangular matrix product on the DMMA pipe (sm_100a).
//
//   Out[b][i][c] = sum_k A[b][i][k] * B[b][k][c]      k <= row-tile end (LOWER)  or  k >= row-tile start (UPPER)
//
// A is the cached triangular factor of one objective (L^-1 row-major for LOWER, L^-T row-major for UPPER,
// zero padded), B is a training-index-major slab of candidates ([k][c], c contiguous).  This is the
// dominant kernel of the posterior: it replaces `np.dot(K, gp._K_inv)` / `dK_T @ gp._K_inv`
// (autooed/mobo/surrogate_model/gp.py:160,202-203) with w = L^-1 k*, v = L^-T w.
//
// Tiling: CTA 128(i) x 128(c), k-step 16, 4-stage cp.async ring, 16 warps (4x4) with 32x32 warp tiles of
// m8n8k4 DMMAs.  A stage rows are padded to 18 doubles and the k index inside a 16-block is permuted
// (step s of lane quad-index q uses k = 4q+s) so that every A fragment is two conflict-free LDS.128;
// the B stage is dense [16][128] with a 32-byte-group XOR swizzle on (k>>2)&3, conflict-free LDS.64.
// Optional fused epilogue: per-column sum of squares over the rows of the tile (sigma^2 = k** - |w|^2).
#pragma once
#include "common.cuh"

namespace aoed {

constexpr int TM = 128;
constexpr int TN = 128;  // column granularity of the slabs (workspace columns are padded to this)
constexpr int TK = 16;
constexpr int TSTAGES = 4;
constexpr int A_LD = 18;
constexpr int A_STAGE = TM * A_LD;
// Two tilings of the same kernel: 128x128 (16 warps, one CTA per SM) and 128x64 (8 warps, two CTAs per SM whose
// barrier bubbles overlap each other).
template <int TNV>
struct TrmmCfg {
    static constexpr int THREADS = TNV * 4;  // (TM/32) x (TNV/32) warps of 32x32
    static constexpr int B_STAGE = TK * TNV;
    static constexpr int STAGE_DOUBLES = A_STAGE + B_STAGE;
    static constexpr size_t SMEM = size_t(TSTAGES) * STAGE_DOUBLES * sizeof(double);
    static constexpr int MIN_CTAS = TNV == 128 ? 1 : 2;
};

struct TrmmArgs {
    const double* A;  // [batch][mTiles*TM][lda]
    int64_t strideA;
    int lda;
    const double* B;  // [batch][>=Kpad rows][ldb]
    int64_t strideB;
    int ldb;
    double* Out;  // [batch][mTiles*TM][ldo] or nullptr
    int64_t strideOut;
    int ldo;
    double* sumsq;  // [batch][mTiles][ldo] or nullptr
    int64_t strideSS;
    int M;       // valid rows
    int Kpad;    // valid k rounded up to TK
    int mTiles;
    int colTiles;
};

template <bool UPPER, int TNV>
__global__ void __launch_bounds__(TrmmCfg<TNV>::THREADS, TrmmCfg<TNV>::MIN_CTAS) trmm_kernel(TrmmArgs p) {
    using Cfg = TrmmCfg<TNV>;
    constexpr int TTHREADS = Cfg::THREADS, STAGE_DOUBLES = Cfg::STAGE_DOUBLES, WN = TNV / 32;
    constexpr int TN = TNV;  // shadows the slab granularity inside the kernel
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp -> 32x32 sub-tile.  Warps w and w+4 share an SM sub-partition; the k-range of a row slice inside the diagonal
    // block grows with wm, so the 8-warp tiling pairs wm with 3-wm on every sub-partition (equal DMMA work per scheduler).
    const int wm = (WN == 4) ? warp / WN : ((warp < 4) ? warp : 7 - warp);
    const int wn = (WN == 4) ? warp % WN : (warp >> 2);
    const int g = lane >> 2, q = lane & 3;
    // Work per CTA grows with the k-range of its row tile.  CTAs are dispatched in linear block order, so the row
    // tile is the SLOW grid index, heaviest first: every heavy tile of the launch starts before any light one and the
    // light tiles fill the tail (longest-processing-time-first).
    const int mt = UPPER ? int(blockIdx.y) : (p.mTiles - 1 - int(blockIdx.y));
    const int ct = int(blockIdx.x) % p.colTiles, batch = int(blockIdx.x) / p.colTiles;
    const double* __restrict__ A = p.A + batch * p.strideA + int64_t(mt) * TM * p.lda;
    const double* __restrict__ B = p.B + batch * p.strideB + int64_t(ct) * TN;
    const int k_begin = UPPER ? mt * TM : 0;
    const int k_end = UPPER ? p.Kpad : min(p.Kpad, (mt + 1) * TM);
    const int KT = (k_end - k_begin) / TK;
    // k16-steps of the diagonal 128-block this warp can skip (its rows only see zeros there)
    const int diag0 = UPPER ? 0 : mt * (TM / TK);  // first diagonal step (step index)
    const int lda = p.lda, ldb = p.ldb;
    // row slices entirely in the zero padding of the last row tile (N not a multiple of 128) do no arithmetic
    const bool rows_live = mt * TM + wm * 32 < p.M;

    auto load_stage = [&](int slot, int kt) {
        double* As = smem + slot * STAGE_DOUBLES;
        double* Bs = As + A_STAGE;
        const int k0 = k_begin + kt * TK;
#pragma unroll
        for (int i = 0; i < TM * 8 / TTHREADS; ++i) {
            int row = (tid >> 3) + (TTHREADS / 8) * i, ch = tid & 7;
            cp_async16(As + row * A_LD + ch * 2, A + int64_t(row) * lda + k0 + ch * 2, true);
        }
        constexpr int CPR = TN / 2;  // 16-byte chunks per B row
#pragma unroll
        for (int i = 0; i < TK * CPR / TTHREADS; ++i) {
            int k = tid / CPR + (TTHREADS / CPR) * i, ch = tid % CPR;
            int phys = (((ch >> 1) ^ ((k >> 2) & 3)) << 1) | (ch & 1);
            cp_async16(Bs + k * TN + phys * 2, B + int64_t(k0 + k) * ldb + ch * 2, true);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < TSTAGES - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<TSTAGES - 2>();
        __syncthreads();
        {
            int nk = kt + TSTAGES - 1;
            if (nk < KT) load_stage(nk % TSTAGES, nk);
            cp_async_commit();
        }
        // triangular skip inside the diagonal block
        const int kd = kt - diag0;  // 0..7 inside the diagonal block
        bool active = rows_live;
        if (UPPER) {
            if (kd < 2 * wm) active = false;  // rows >= mt*128 + wm*32 need k >= that
        } else {
            if (kd > 2 * wm + 1) active = false;  // rows < mt*128 + wm*32 + 32 need k <= that
        }
        if (active) {
            const double* As = smem + (kt % TSTAGES) * STAGE_DOUBLES;
            const double* Bs = As + A_STAGE;
            double a[4][4], b[4][4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                const double2* pa = reinterpret_cast<const double2*>(As + (wm * 32 + mb * 8 + g) * A_LD + 4 * q);
                double2 v0 = pa[0], v1 = pa[1];
                a[mb][0] = v0.x; a[mb][1] = v0.y; a[mb][2] = v1.x; a[mb][3] = v1.y;
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const int n = wn * 32 + nb * 8 + g;
                const int col = (((n >> 2) ^ q) << 2) | (n & 3);
#pragma unroll
                for (int s = 0; s < 4; ++s) b[nb][s] = Bs[(4 * q + s) * TN + col];
            }
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb][s], b[nb][s]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    if (p.Out != nullptr) {
        double* __restrict__ Out = p.Out + batch * p.strideOut;
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
            const int row = mt * TM + wm * 32 + mb * 8 + g;
            if (row < p.M) {
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const int col = ct * TN + wn * 32 + nb * 8 + 2 * q;
                    *reinterpret_cast<double2*>(Out + int64_t(row) * p.ldo + col) =
                        make_double2(acc[mb][nb][0], acc[mb][nb][1]);
                }
            }
        }
    }
    if (p.sumsq != nullptr) {
        double* red = smem;  // [4][128]
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double ss = 0.0;
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) ss = fma(acc[mb][nb][e], acc[mb][nb][e], ss);
                ss += __shfl_xor_sync(0xffffffffu, ss, 4);
                ss += __shfl_xor_sync(0xffffffffu, ss, 8);
                ss += __shfl_xor_sync(0xffffffffu, ss, 16);
                if (g == 0) red[wm * TN + wn * 32 + nb * 8 + 2 * q + e] = ss;
            }
        }
        __syncthreads();
        if (tid < TN) {
            double s = (red[tid] + red[TN + tid]) + (red[2 * TN + tid] + red[3 * TN + tid]);
            p.sumsq[batch * p.strideSS + int64_t(mt) * p.ldo + ct * TN + tid] = s;
        }
    }
}

}  // namespace aoed
