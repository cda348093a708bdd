// Multi-objective selection kernels (sm_100a): Pareto mask / rank and greedy hypervolume improvement.
// Contracts in include/aoed_b200.h.  Replaces autooed/utils/pareto.py:27-62 and
// autooed/mobo/selection/hvi.py:16-47 (+ the pymoo hypervolume it calls).
//
// Hypervolume improvement is computed DIRECTLY as the exclusive volume of box(y, ref) not dominated by the
// front, by a dimension sweep: slabs along the last objective, recursing down to a 2-D staircase sweep over
// the front sorted by its first objective.  A dominated candidate yields exactly 0, as in the reference where
// pymoo's non-dominated pre-filter makes HV(P u {y}) - HV(P) vanish.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/aoed_b200.h"
#include "common.cuh"
#include "rank.cuh"

extern "C" int aoed_detail_fail(int code, const char* msg);  // defined in gp.cu (shared last-error slot)

namespace {

using aoed::RS_MAX_N;

int mfail(int code, const std::string& msg) { return aoed_detail_fail(code, msg.c_str()); }
#define MCU(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return mfail(AOED_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + \
                                            __FILE__ + ":" + std::to_string(__LINE__));             \
    } while (0)

#define MCHK(expr)                      \
    do {                                \
        int rc_ = (expr);               \
        if (rc_ != AOED_OK) return rc_; \
    } while (0)

constexpr int MOO_MAX_M = 8;

// ---- Pareto ------------------------------------------------------------------------------------------
// YT: [m][n] objective-major.  alive == nullptr -> all points take part.  dominated[i] = 1 iff some alive j
// has all(Y_j <= Y_i) and any(Y_j < Y_i)   (utils/pareto.py:39).
template <int M>
__global__ void __launch_bounds__(256) dominated_kernel(const double* __restrict__ YT, int64_t n, int64_t row0,
                                                        int64_t nrows, const uint8_t* __restrict__ alive,
                                                        uint8_t* __restrict__ dominated) {
    __shared__ double tile[M][256];
    __shared__ uint8_t talive[256];
    const int64_t li = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;  // row of this shard
    const int64_t i = row0 + li;
    double yi[M];
#pragma unroll
    for (int k = 0; k < M; ++k) yi[k] = (li < nrows) ? YT[k * n + i] : 0.0;
    bool dom = false;
    for (int64_t j0 = 0; j0 < n; j0 += 256) {
        const int64_t j = j0 + threadIdx.x;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < M; ++k) tile[k][threadIdx.x] = (j < n) ? YT[k * n + j] : DBL_MAX;
        talive[threadIdx.x] = (j < n) ? (alive ? alive[j] : 1) : 0;
        __syncthreads();
        const int lim = (n - j0 < 256) ? int(n - j0) : 256;
        for (int t = 0; t < lim; ++t) {
            bool le = true, lt = false;
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const double v = tile[k][t];
                le = le && (v <= yi[k]);
                lt = lt || (v < yi[k]);
            }
            dom = dom || (le && lt && talive[t]);
        }
    }
    if (li < nrows) dominated[li] = dom ? 1 : 0;
}

template <int M>
void launch_dominated_m(const double* YT, int64_t n, int64_t row0, int64_t nrows, const uint8_t* alive, uint8_t* dom,
                        cudaStream_t s) {
    dominated_kernel<M><<<unsigned((nrows + 255) / 256), 256, 0, s>>>(YT, n, row0, nrows, alive, dom);
}

int launch_dominated(int m, const double* YT, int64_t n, const uint8_t* alive, uint8_t* dom, cudaStream_t s,
                     int64_t row0 = 0, int64_t nrows = -1) {
    if (nrows < 0) nrows = n;
    if (nrows == 0) return AOED_OK;
    switch (m) {
        case 1: launch_dominated_m<1>(YT, n, row0, nrows, alive, dom, s); break;
        case 2: launch_dominated_m<2>(YT, n, row0, nrows, alive, dom, s); break;
        case 3: launch_dominated_m<3>(YT, n, row0, nrows, alive, dom, s); break;
        case 4: launch_dominated_m<4>(YT, n, row0, nrows, alive, dom, s); break;
        case 5: launch_dominated_m<5>(YT, n, row0, nrows, alive, dom, s); break;
        case 6: launch_dominated_m<6>(YT, n, row0, nrows, alive, dom, s); break;
        case 7: launch_dominated_m<7>(YT, n, row0, nrows, alive, dom, s); break;
        case 8: launch_dominated_m<8>(YT, n, row0, nrows, alive, dom, s); break;
        default: return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    }
    MCU(cudaGetLastError());
    return AOED_OK;
}

// rank pass: alive & not dominated -> rank = r, alive = 0; counts how many remain alive
__global__ void rank_update_kernel(int64_t n, const uint8_t* dominated, uint8_t* alive, int32_t* rank, int r,
                                   unsigned long long* remaining) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n || !alive[i]) return;
    if (!dominated[i]) {
        rank[i] = r;
        alive[i] = 0;
    } else {
        atomicAdd(remaining, 1ull);
    }
}

int launch_rank_small(int m, const double* YT, int n, int32_t* rank, int* cnt_scratch, cudaStream_t s) {
#define RS_CASE(MM)                                                                       \
    case MM:                                                                              \
        MCU(aoed::launch_rank_small_m<MM>(YT, n, rank, cnt_scratch, /*need=*/n, s));      \
        break;
    switch (m) {
        RS_CASE(1) RS_CASE(2) RS_CASE(3) RS_CASE(4) RS_CASE(5) RS_CASE(6) RS_CASE(7) RS_CASE(8)
        default: return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    }
#undef RS_CASE
    return AOED_OK;
}

// ---- exclusive hypervolume -----------------------------------------------------------------------------
// front: [k][M] sorted ascending by objective 0 (shared memory).  lev[] holds, for objectives >= D, the slab
// level: a front point is active iff p[t] <= lev[t] for all t in [D, M).
template <int M>
__device__ __forceinline__ bool is_active(const double* p, const double* lev, int D) {
    bool a = true;
#pragma unroll
    for (int t = 2; t < M; ++t)
        if (t >= D) a = a && (p[t] <= lev[t]);
    return a;
}

template <int M>
__device__ double excl2d(const double* y, const double* front, int k, const double* ref, const double* lev) {
    const double a = y[0], b = y[1];
    double hmin = ref[1], xprev = a, area = 0.0;
    for (int i = 0; i < k; ++i) {
        const double* p = front + i * M;
        if (!is_active<M>(p, lev, 2)) continue;
        const double px = p[0];
        if (px >= ref[0]) break;
        if (px > a) {
            const double h = hmin - b;
            if (h <= 0.0) return area;
            area = fma(px - xprev, h, area);
            xprev = px;
        }
        hmin = fmin(hmin, p[1]);
    }
    const double h = hmin - b;
    if (h > 0.0) area = fma(ref[0] - xprev, h, area);
    return area;
}

template <int M, int D>
struct Excl {
    // exclusive volume in objectives [0, D) of y, slabbing along objective D-1
    static __device__ double run(const double* y, const double* front, int k, const double* ref, double* lev) {
        double vol = 0.0;
        double cur = y[D - 1];
        const double top = ref[D - 1];
        while (cur < top) {
            // next slab boundary: smallest front coordinate above cur (among points active in higher dims)
            double nxt = top;
            for (int i = 0; i < k; ++i) {
                const double* p = front + i * M;
                const double v = p[D - 1];
                if (v > cur && v < nxt && is_active<M>(p, lev, D)) nxt = v;
            }
            lev[D - 1] = cur;
            const double sub = Excl<M, D - 1>::run(y, front, k, ref, lev);
            if (sub <= 0.0) break;  // the cross-section only shrinks as the level rises
            vol = fma(nxt - cur, sub, vol);
            cur = nxt;
        }
        return vol;
    }
};
template <int M>
struct Excl<M, 2> {
    static __device__ double run(const double* y, const double* front, int k, const double* ref, double* lev) {
        return excl2d<M>(y, front, k, ref, lev);
    }
};
template <int M>
struct Excl<M, 1> {
    static __device__ double run(const double* y, const double*, int, const double* ref, double*) { return ref[0] - y[0]; }
};

// One thread per candidate.  contrib[i] = exclusive hypervolume (0 for excluded / out-of-box candidates).
template <int M>
__global__ void __launch_bounds__(128) hvc_kernel(const double* __restrict__ YcT, int64_t n, const double* __restrict__ front,
                                                  int k, const double* __restrict__ ref, const uint8_t* __restrict__ excluded,
                                                  double* __restrict__ contrib) {
    extern __shared__ double sfront[];
    __shared__ double sref[M];
    for (int i = threadIdx.x; i < k * M; i += blockDim.x) sfront[i] = front[i];
    if (threadIdx.x < M) sref[threadIdx.x] = ref[threadIdx.x];
    __syncthreads();
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double y[M], lev[M];
    bool inside = true;
#pragma unroll
    for (int t = 0; t < M; ++t) {
        y[t] = YcT[t * n + i];
        lev[t] = y[t];
        inside = inside && (y[t] <= sref[t]);
    }
    double v = 0.0;
    if (inside && !(excluded && excluded[i])) {
        if (M == 1) {
            double best = sref[0];
            for (int j = 0; j < k; ++j) best = fmin(best, sfront[j]);
            v = fmax(0.0, best - y[0]);
        } else {
            v = Excl<M, M>::run(y, sfront, k, sref, lev);
        }
    }
    contrib[i] = v;
}

int launch_hvc(int m, const double* YcT, int64_t n, const double* front, int k, const double* ref,
               const uint8_t* excluded, double* contrib, cudaStream_t s) {
    const unsigned grid = unsigned((n + 127) / 128);
#define HVC_CASE(MM)                                                                                            \
    case MM: {                                                                                                  \
        const size_t sm = size_t(std::max(k, 1)) * MM * sizeof(double);                                         \
        if (sm > 48 * 1024) MCU(cudaFuncSetAttribute(hvc_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm))); \
        hvc_kernel<MM><<<grid, 128, sm, s>>>(YcT, n, front, k, ref, excluded, contrib);                         \
        break;                                                                                                  \
    }
    switch (m) {
        HVC_CASE(1)
        HVC_CASE(2)
        HVC_CASE(3)
        HVC_CASE(4)
        HVC_CASE(5)
        HVC_CASE(6)
        HVC_CASE(7)
        HVC_CASE(8)
        default: return mfail(AOED_ERR_UNSUPPORTED, "hypervolume improvement supports n_obj <= 8");
    }
#undef HVC_CASE
    MCU(cudaGetLastError());
    return AOED_OK;
}

// ---- device-resident greedy rounds ---------------------------------------------------------------------------------
// One kernel per greedy round of hvi.py:24-42, no host synchronisation between rounds: every thread scores candidates
// (grid-stride), the two best (value, lowest index on ties — hvi.py:35-37 keeps the first maximum) are reduced per
// block, the last block to finish reduces the block results, records the round, marks the winner as taken and
// inserts its objective vector into the front (kept sorted by objective 0 in device memory, appended after equal keys
// exactly like the stable sort of the host path it replaces).  When no candidate has a positive contribution the loop
// parks (`stopped`): the caller draws the reference's random pick (hvi.py:38-39) with the host RNG and calls again.
struct Top2 {
    double v1, v2;
    long long i1, i2;
};
__device__ __forceinline__ bool top_better(double va, long long ia, double vb, long long ib) {
    return va > vb || (va == vb && ia < ib);
}
__device__ __forceinline__ void top_push(Top2& t, double v, long long i) {
    if (top_better(v, i, t.v1, t.i1)) {
        t.v2 = t.v1; t.i2 = t.i1; t.v1 = v; t.i1 = i;
    } else if (top_better(v, i, t.v2, t.i2)) {
        t.v2 = v; t.i2 = i;
    }
}

struct HviState {
    int k;        // current front size
    int stopped;  // 1: a round found no positive contribution
    int round;    // rounds recorded so far
    unsigned int done_blocks;
};

struct HviArgs {
    const double* YcT;  // [M][n]
    long long n;
    double* front;      // [front_cap][M], sorted by objective 0
    const double* ref;
    uint8_t* excluded;
    HviState* st;
    Top2* partial;      // [gridDim.x]
    long long* out_i1;  // per round: winner, its contribution, runner-up contribution and index
    double* out_v1;
    double* out_v2;
    long long* out_i2;
};

constexpr int HVI_THREADS = 128;
constexpr int HVI_UNROLL = 4;

__device__ __forceinline__ Top2 block_top2(Top2 t, Top2* sm) {
    sm[threadIdx.x] = t;
    __syncthreads();
    for (int off = HVI_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            Top2 a = sm[threadIdx.x];
            const Top2 o = sm[threadIdx.x + off];
            top_push(a, o.v1, o.i1);
            top_push(a, o.v2, o.i2);
            sm[threadIdx.x] = a;
        }
        __syncthreads();
    }
    return sm[0];
}

template <int M>
__global__ void __launch_bounds__(HVI_THREADS) hvi_round_kernel(HviArgs a) {
    extern __shared__ double sfront[];
    __shared__ double sref[M];
    __shared__ Top2 stop[HVI_THREADS];
    __shared__ int s_last, s_pos;
    const int k = a.st->k;
    if (a.st->stopped) return;
    for (int i = threadIdx.x; i < k * M; i += HVI_THREADS) sfront[i] = a.front[i];
    if (threadIdx.x < M) sref[threadIdx.x] = a.ref[threadIdx.x];
    __syncthreads();
    Top2 t;
    t.v1 = t.v2 = -1.0;
    t.i1 = t.i2 = 0x7fffffffffffffffLL;
    // HVI_UNROLL candidates per thread and step: their loads (8 m bytes + the taken flag each) are issued together — the scan is
    // bound by memory latency, not by the few dozen instructions of a dominated candidate's sweep
    const long long stride = (long long)gridDim.x * HVI_THREADS;
    for (long long i0 = blockIdx.x * (long long)HVI_THREADS + threadIdx.x; i0 < a.n; i0 += HVI_UNROLL * stride) {
        double yy[HVI_UNROLL][M];
        uint8_t ex[HVI_UNROLL];
#pragma unroll
        for (int u = 0; u < HVI_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            ex[u] = 1;
            if (i < a.n) {
                ex[u] = a.excluded[i];
#pragma unroll
                for (int c = 0; c < M; ++c) yy[u][c] = a.YcT[c * a.n + i];
            }
        }
#pragma unroll
        for (int u = 0; u < HVI_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i >= a.n) break;
            double y[M], lev[M];
            bool inside = true;
#pragma unroll
            for (int c = 0; c < M; ++c) {
                y[c] = yy[u][c];
                lev[c] = y[c];
                inside = inside && (y[c] <= sref[c]);
            }
            double v = 0.0;
            if (inside && !ex[u]) {
                if (M == 1) {
                    double best = sref[0];
                    for (int j = 0; j < k; ++j) best = fmin(best, sfront[j]);
                    v = fmax(0.0, best - y[0]);
                } else {
                    v = Excl<M, M>::run(y, sfront, k, sref, lev);
                }
                // contributions only shrink as the front grows: a candidate at zero stays at zero and is skipped from now on
                if (!(v > 0.0)) a.excluded[i] = 2;
            }
            top_push(t, v, i);
        }
    }
    t = block_top2(t, stop);
    if (threadIdx.x == 0) {
        a.partial[blockIdx.x] = t;
        __threadfence();
        s_last = (atomicAdd(&a.st->done_blocks, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last block: final reduction, bookkeeping, front insertion ----
    __threadfence();
    t.v1 = t.v2 = -1.0;
    t.i1 = t.i2 = 0x7fffffffffffffffLL;
    const volatile Top2* part = a.partial;
    for (int b = threadIdx.x; b < int(gridDim.x); b += HVI_THREADS) {
        top_push(t, part[b].v1, part[b].i1);
        top_push(t, part[b].v2, part[b].i2);
    }
    __syncthreads();
    t = block_top2(t, stop);
    const bool found = t.v1 > 0.0;
    const int r = a.st->round;
    double y[M];
    if (found) {
#pragma unroll
        for (int c = 0; c < M; ++c) y[c] = a.YcT[c * a.n + t.i1];
    }
    if (threadIdx.x == 0) {
        s_pos = 0;
        a.out_i1[r] = found ? t.i1 : -1;
        a.out_v1[r] = found ? t.v1 : 0.0;
        a.out_v2[r] = t.v2;
        a.out_i2[r] = (t.v2 >= 0.0) ? t.i2 : -1;
    }
    __syncthreads();
    if (found) {
        int cnt = 0;  // entries with key <= y0 stay in front of the new point (stable insertion)
        for (int j = threadIdx.x; j < k; j += HVI_THREADS) cnt += (sfront[j * M] <= y[0]) ? 1 : 0;
        if (cnt) atomicAdd(&s_pos, cnt);
        __syncthreads();
        const int pos = s_pos;
        for (int e = pos * M + threadIdx.x; e < k * M; e += HVI_THREADS) a.front[e + M] = sfront[e];
        if (threadIdx.x < M) a.front[pos * M + threadIdx.x] = y[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        if (found) {
            a.excluded[t.i1] = 1;
            a.st->k = k + 1;
        } else {
            a.st->stopped = 1;
        }
        a.st->round = r + 1;
        a.st->done_blocks = 0;
    }
}

int launch_hvi_round(int m, const HviArgs& a, int grid, size_t smem, cudaStream_t s) {
#define HVR_CASE(MM)                                                                                                      \
    case MM:                                                                                                              \
        if (smem > 48 * 1024) MCU(cudaFuncSetAttribute(hvi_round_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
        hvi_round_kernel<MM><<<grid, HVI_THREADS, smem, s>>>(a);                                                          \
        break;
    switch (m) {
        HVR_CASE(1) HVR_CASE(2) HVR_CASE(3) HVR_CASE(4) HVR_CASE(5) HVR_CASE(6) HVR_CASE(7) HVR_CASE(8)
        default: return mfail(AOED_ERR_UNSUPPORTED, "hypervolume improvement supports n_obj <= 8");
    }
#undef HVR_CASE
    MCU(cudaGetLastError());
    return AOED_OK;
}

// Y [n][m] row-major -> YT [m][n] (objective-major), on the device
__global__ void to_objective_major_kernel(const double* __restrict__ Y, long long n, int m, double* __restrict__ YT) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < m; ++k) YT[k * n + i] = Y[i * m + k];
}

// rows of `idx` gathered into an objective-major array: out[k][j] = YT[k][idx[j]]
__global__ void gather_columns_kernel(const double* __restrict__ YT, long long n, int m, const long long* __restrict__ idx,
                                      long long cnt, double* __restrict__ out) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= cnt) return;
    const long long i = idx[j];
    for (int k = 0; k < m; ++k) out[k * cnt + j] = YT[k * n + i];
}

// dominated[r] = 1 iff some column point c has all(C_c <= R_r) and any(C_c < R_r): rows and columns are different sets
template <int M>
__global__ void __launch_bounds__(256) dominated_by_kernel(const double* __restrict__ RT, long long nr_total, long long row0,
                                                           long long nrows, const double* __restrict__ CT, long long nc,
                                                           uint8_t* __restrict__ dominated) {
    __shared__ double tile[M][256];
    const long long li = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double yi[M];
#pragma unroll
    for (int k = 0; k < M; ++k) yi[k] = (li < nrows) ? RT[k * nr_total + row0 + li] : 0.0;
    bool dom = false;
    for (long long j0 = 0; j0 < nc; j0 += 256) {
        const long long j = j0 + threadIdx.x;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < M; ++k) tile[k][threadIdx.x] = (j < nc) ? CT[k * nc + j] : DBL_MAX;
        __syncthreads();
        const int lim = (nc - j0 < 256) ? int(nc - j0) : 256;
        for (int t = 0; t < lim; ++t) {
            bool le = true, lt = false;
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const double v = tile[k][t];
                le = le && (v <= yi[k]);
                lt = lt || (v < yi[k]);
            }
            dom = dom || (le && lt);
        }
    }
    if (li < nrows) dominated[li] = dom ? 1 : 0;
}

int launch_dominated_by(int m, const double* RT, long long nr_total, long long row0, long long nrows, const double* CT,
                        long long nc, uint8_t* dom, cudaStream_t s) {
    if (nrows == 0) return AOED_OK;
    const unsigned grid = unsigned((nrows + 255) / 256);
#define DB_CASE(MM) \
    case MM: dominated_by_kernel<MM><<<grid, 256, 0, s>>>(RT, nr_total, row0, nrows, CT, nc, dom); break;
    switch (m) {
        DB_CASE(1) DB_CASE(2) DB_CASE(3) DB_CASE(4) DB_CASE(5) DB_CASE(6) DB_CASE(7) DB_CASE(8)
        default: return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    }
#undef DB_CASE
    MCU(cudaGetLastError());
    return AOED_OK;
}

__global__ void flip_kernel(const uint8_t* __restrict__ dominated, long long n, uint8_t* __restrict__ mask) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) mask[i] = dominated[i] ? 0 : 1;
}

// indices i in [0, n) with flag[i] == want, in unspecified order; *count must be 0 on entry
__global__ void compact_kernel(const uint8_t* __restrict__ flag, long long n, uint8_t want, long long stride,
                               long long* __restrict__ out, unsigned long long* count) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long i = t * stride;
    if (i >= n) return;
    if (flag == nullptr || flag[i] == want) out[atomicAdd(count, 1ull)] = i;
}

__global__ void scatter_mask_kernel(const long long* __restrict__ idx, const uint8_t* __restrict__ dominated, long long cnt,
                                    long long row_lo, long long row_hi, uint8_t* __restrict__ mask) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= cnt) return;
    const long long i = idx[j];
    if (i >= row_lo && i < row_hi && !dominated[j]) mask[i - row_lo] = 1;
}

// ---- cached device workspace (one per device; calls are serialised by g_moo_mu) ------------------------------------
struct MooWs {
    static constexpr int SLOTS = 16;
    void* p[SLOTS] = {};
    size_t cap[SLOTS] = {};
};
std::mutex g_moo_mu;
MooWs g_moo_ws[64];

template <typename T>
int ws_get(MooWs& w, int slot, size_t count, T** out) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if (w.cap[slot] < bytes) {
        if (w.p[slot]) cudaFree(w.p[slot]);
        w.p[slot] = nullptr;
        w.cap[slot] = 0;
        const size_t want = bytes + bytes / 4;
        MCU(cudaMalloc(&w.p[slot], want));
        w.cap[slot] = want;
    }
    *out = static_cast<T*>(w.p[slot]);
    return AOED_OK;
}
enum { WS_Y = 0, WS_YT, WS_FRONT, WS_REF, WS_CONTRIB, WS_EXCL, WS_STATE, WS_PART, WS_OUT, WS_DOM, WS_IDX, WS_SUB, WS_DOM2, WS_CNT, WS_MASK };

// uploads row-major Y and leaves the objective-major copy in slot WS_YT
int upload_objective_major(MooWs& w, const double* h_Y, int64_t n, int m, double** dYT) {
    double* dY;
    MCHK(ws_get(w, WS_Y, size_t(n) * m, &dY));
    MCHK(ws_get(w, WS_YT, size_t(n) * m, dYT));
    MCU(cudaMemcpy(dY, h_Y, size_t(n) * m * sizeof(double), cudaMemcpyHostToDevice));
    to_objective_major_kernel<<<unsigned((n + 255) / 256), 256>>>(dY, n, m, *dYT);
    MCU(cudaGetLastError());
    return AOED_OK;
}

std::vector<double> sort_front(const double* front, int64_t k, int m) {
    std::vector<int64_t> idx(k);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return front[a * m] < front[b * m]; });
    std::vector<double> out(size_t(k) * m);
    for (int64_t i = 0; i < k; ++i)
        for (int t = 0; t < m; ++t) out[size_t(i) * m + t] = front[idx[i] * m + t];
    return out;
}

// ---- host-side exact hypervolume in the reference's DIFFERENCE form (near-tie re-decision) --------------------------
// Volume of the union of the boxes [0, q]: slicing along the last coordinate, the same sequence of floating-point
// operations as the test suite's CPU restatement of pymoo's exact indicator, so that HV(P u {y}) - HV(P) (hvi.py:28-34)
// orders near-ties the way that formulation does.  volatile temporaries keep the compiler from contracting a*b+c.
double hv_boxes(std::vector<const double*> q, int m) {
    const size_t k = q.size();
    if (k == 0) return 0.0;
    if (m == 1) {
        double best = q[0][0];
        for (size_t i = 1; i < k; ++i) best = std::max(best, q[i][0]);
        return best;
    }
    if (m == 2) {
        std::stable_sort(q.begin(), q.end(), [](const double* a, const double* b) { return a[0] > b[0]; });
        double hv = 0.0, best1 = 0.0;
        for (size_t i = 0; i < k; ++i)
            if (q[i][1] > best1) {
                volatile double h = q[i][1] - best1;
                volatile double area = q[i][0] * h;
                hv = hv + area;
                best1 = q[i][1];
            }
        return hv;
    }
    std::stable_sort(q.begin(), q.end(), [m](const double* a, const double* b) { return a[m - 1] > b[m - 1]; });
    double hv = 0.0;
    for (size_t i = 0; i < k; ++i) {
        const double nxt = (i + 1 < k) ? q[i + 1][m - 1] : 0.0;
        volatile double depth = q[i][m - 1] - nxt;
        if (depth > 0) {
            volatile double slab = depth * hv_boxes(std::vector<const double*>(q.begin(), q.begin() + i + 1), m - 1);
            hv = hv + slab;
        }
    }
    return hv;
}

// pymoo `Hypervolume(ref_point).calc(F)` for minimisation: only the first non-dominated front of the rows is measured
// (equal rows do not dominate each other, original order kept), rows with a coordinate above ref are dropped
double host_hypervolume(const std::vector<double>& F, const double* extra, int m, const double* ref) {
    const size_t k = F.size() / m + (extra ? 1 : 0);
    auto row = [&](size_t i) { return (i < F.size() / m) ? &F[i * m] : extra; };
    std::vector<double> Q;
    Q.reserve(k * m);
    size_t kept = 0;
    for (size_t i = 0; i < k; ++i) {
        const double* f = row(i);
        bool dominated = false;
        for (size_t j = 0; j < k && !dominated; ++j) {
            const double* g = row(j);
            bool le = true, lt = false;
            for (int t = 0; t < m; ++t) {
                le = le && (g[t] <= f[t]);
                lt = lt || (g[t] < f[t]);
            }
            dominated = le && lt;
        }
        if (dominated) continue;
        bool in = true;
        for (int t = 0; t < m; ++t) in = in && (f[t] <= ref[t]);
        if (!in) continue;
        for (int t = 0; t < m; ++t) Q.push_back(ref[t] - f[t]);
        ++kept;
    }
    std::vector<const double*> rows(kept);
    for (size_t i = 0; i < kept; ++i) rows[i] = &Q[i * m];
    return hv_boxes(rows, m);
}

// Non-dominated test of the rows [row_lo, row_hi) of YT (objective-major, n points) against ALL n points, result in
// dMask (1 = non-dominated).  Small sets: one tiled all-pairs pass.  Large sets (> PF_DIRECT rows): a two-stage filter —
// (1) the skyline S0 of a strided sample of PF_SAMPLE points, (2) every point tested against S0 only: what survives is a
// small superset of the skyline (a dominated point is always dominated by a skyline point, and skyline points are never
// dropped), (3) exact all-pairs pass among the survivors — O(n |S0| + |R|^2) instead of O(n^2), identical result.
constexpr int64_t PF_DIRECT = 16384, PF_SAMPLE = 4096;

int pareto_mask_device(MooWs& w, int m, const double* dYT, int64_t n, int64_t row_lo, int64_t row_hi, uint8_t* dMask) {
    const int64_t nrows = row_hi - row_lo;
    const char* force = getenv("AOED_PARETO_PATH");  // "direct" / "filter" force a path (tests)
    const bool filter = force ? strcmp(force, "filter") == 0 : n > PF_DIRECT;
    uint8_t* dDom;
    MCHK(ws_get(w, WS_DOM, size_t(n), &dDom));
    if (!filter) {
        int rc = launch_dominated(m, dYT, n, nullptr, dDom, nullptr, row_lo, nrows);
        if (rc != AOED_OK) return rc;
        flip_kernel<<<unsigned((nrows + 255) / 256), 256>>>(dDom, nrows, dMask);
        MCU(cudaGetLastError());
        return AOED_OK;
    }
    long long *dIdx, *dIdx2;
    unsigned long long* dCnt;
    double *dSub, *dSub2;
    uint8_t* dDom2;
    MCHK(ws_get(w, WS_IDX, size_t(2 * n), &dIdx));
    dIdx2 = dIdx + n;
    MCHK(ws_get(w, WS_CNT, 4, &dCnt));
    MCU(cudaMemset(dCnt, 0, 4 * sizeof(unsigned long long)));
    // (1) strided sample -> its skyline S0
    const int64_t stride = std::max<int64_t>(1, n / PF_SAMPLE);
    const int64_t ns = (n + stride - 1) / stride;
    compact_kernel<<<unsigned((ns + 255) / 256), 256>>>(nullptr, n, 0, stride, dIdx, dCnt);
    MCHK(ws_get(w, WS_SUB, size_t(2 * ns) * m, &dSub));
    dSub2 = dSub + ns * m;
    gather_columns_kernel<<<unsigned((ns + 255) / 256), 256>>>(dYT, n, m, dIdx, ns, dSub);
    MCHK(ws_get(w, WS_DOM2, size_t(ns), &dDom2));
    int rc = launch_dominated_by(m, dSub, ns, 0, ns, dSub, ns, dDom2, nullptr);
    if (rc != AOED_OK) return rc;
    compact_kernel<<<unsigned((ns + 255) / 256), 256>>>(dDom2, ns, 0, 1, dIdx2, dCnt + 1);
    unsigned long long s0 = 0;
    MCU(cudaMemcpy(&s0, dCnt + 1, sizeof(s0), cudaMemcpyDeviceToHost));
    gather_columns_kernel<<<unsigned((s0 + 255) / 256), 256>>>(dSub, ns, m, dIdx2, (long long)s0, dSub2);
    // (2) all points against S0
    rc = launch_dominated_by(m, dYT, n, 0, n, dSub2, (long long)s0, dDom, nullptr);
    if (rc != AOED_OK) return rc;
    compact_kernel<<<unsigned((n + 255) / 256), 256>>>(dDom, n, 0, 1, dIdx, dCnt + 2);
    unsigned long long r = 0;
    MCU(cudaMemcpy(&r, dCnt + 2, sizeof(r), cudaMemcpyDeviceToHost));
    // (3) exact pass among the survivors
    MCHK(ws_get(w, WS_SUB, size_t(r) * m + size_t(2 * ns) * m, &dSub));  // may move: S0 is not needed any more
    gather_columns_kernel<<<unsigned((r + 255) / 256), 256>>>(dYT, n, m, dIdx, (long long)r, dSub);
    MCHK(ws_get(w, WS_DOM2, size_t(std::max<unsigned long long>(r, ns)), &dDom2));
    rc = launch_dominated_by(m, dSub, (long long)r, 0, (long long)r, dSub, (long long)r, dDom2, nullptr);
    if (rc != AOED_OK) return rc;
    MCU(cudaMemset(dMask, 0, size_t(nrows)));
    scatter_mask_kernel<<<unsigned((r + 255) / 256), 256>>>(dIdx, dDom2, (long long)r, row_lo, row_hi, dMask);
    MCU(cudaGetLastError());
    return AOED_OK;
}

}  // namespace

extern "C" {

int aoed_pareto_mask(int device, int64_t n, int m, const double* h_Y, uint8_t* h_mask) {
    return aoed_pareto_mask_rows(device, n, m, h_Y, 0, n, h_mask);
}

int aoed_pareto_mask_rows(int device, int64_t n, int m, const double* h_Y, int64_t row_lo, int64_t row_hi,
                          uint8_t* h_mask) {
    if (n < 0 || m < 1 || row_lo < 0 || row_hi < row_lo || row_hi > n || (n > 0 && !h_Y) || (row_hi > row_lo && !h_mask))
        return mfail(AOED_ERR_INVALID, "bad argument");
    const int64_t nrows = row_hi - row_lo;
    if (nrows == 0) return AOED_OK;
    if (m > MOO_MAX_M) return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev || device < 0) return mfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    MCU(guard_.err);
    std::lock_guard<std::mutex> lock(g_moo_mu);
    MooWs& w = g_moo_ws[device & 63];
    double* dYT;
    int rc = upload_objective_major(w, h_Y, n, m, &dYT);
    if (rc != AOED_OK) return rc;
    uint8_t* dMask;
    MCHK(ws_get(w, WS_MASK, size_t(nrows), &dMask));
    rc = pareto_mask_device(w, m, dYT, n, row_lo, row_hi, dMask);
    if (rc != AOED_OK) return rc;
    MCU(cudaMemcpy(h_mask, dMask, size_t(nrows), cudaMemcpyDeviceToHost));
    return AOED_OK;
}

int aoed_pareto_rank(int device, int64_t n, int m, const double* h_Y, int32_t* h_rank) {
    if (n < 0 || m < 1 || (n > 0 && (!h_Y || !h_rank))) return mfail(AOED_ERR_INVALID, "bad argument");
    if (n == 0) return AOED_OK;
    if (m > MOO_MAX_M) return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev || device < 0) return mfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    MCU(guard_.err);
    std::lock_guard<std::mutex> lock(g_moo_mu);
    MooWs& w = g_moo_ws[device & 63];
    double* dYT;
    int rc = upload_objective_major(w, h_Y, n, m, &dYT);
    if (rc != AOED_OK) return rc;
    const char* force = getenv("AOED_RANK_PATH");  // "multi" forces the multi-launch path (tests)
    if (!(force && strcmp(force, "multi") == 0) && n <= RS_MAX_N &&
        aoed::rank_small_smem(m, int(n)) <= 200 * 1024) {
        // one launch: the call is latency bound (it runs once per NSGA-II generation of the host loop)
        int32_t* drank;
        MCHK(ws_get(w, WS_IDX, size_t(n) + aoed::rank_scratch_ints(int(n)), &drank));
        rc = launch_rank_small(m, dYT, int(n), drank, drank + n, nullptr);
        if (rc != AOED_OK) return rc;
        MCU(cudaMemcpy(h_rank, drank, size_t(n) * sizeof(int32_t), cudaMemcpyDeviceToHost));
        return AOED_OK;
    }
    uint8_t *dDom, *dAlive;
    int32_t* dRank;
    unsigned long long* dRem;
    MCHK(ws_get(w, WS_DOM, size_t(n), &dDom));
    MCHK(ws_get(w, WS_EXCL, size_t(n), &dAlive));
    MCHK(ws_get(w, WS_IDX, size_t(n), &dRank));
    MCHK(ws_get(w, WS_CNT, 4, &dRem));
    MCU(cudaMemset(dAlive, 1, size_t(n)));
    MCU(cudaMemset(dRank, 0xff, size_t(n) * sizeof(int32_t)));
    for (int r = 0;; ++r) {
        rc = launch_dominated(m, dYT, n, dAlive, dDom, nullptr);
        if (rc != AOED_OK) return rc;
        MCU(cudaMemset(dRem, 0, sizeof(unsigned long long)));
        rank_update_kernel<<<unsigned((n + 255) / 256), 256>>>(n, dDom, dAlive, dRank, r, dRem);
        MCU(cudaGetLastError());
        unsigned long long rem = 0;
        MCU(cudaMemcpy(&rem, dRem, sizeof(rem), cudaMemcpyDeviceToHost));
        if (rem == 0) break;
        if (r > n) return mfail(AOED_ERR_CUDA, "pareto_rank did not terminate");
    }
    MCU(cudaMemcpy(h_rank, dRank, size_t(n) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return AOED_OK;
}

int aoed_hv_contributions(int device, int64_t n_cand, int m, const double* h_Yc, int64_t n_front,
                          const double* h_front, const double* h_ref, double* h_contrib) {
    if (n_cand < 0 || m < 1 || n_front < 0 || !h_ref || (n_cand > 0 && (!h_Yc || !h_contrib)) || (n_front > 0 && !h_front))
        return mfail(AOED_ERR_INVALID, "bad argument");
    if (n_cand == 0) return AOED_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev || device < 0) return mfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    MCU(guard_.err);
    std::lock_guard<std::mutex> lock(g_moo_mu);
    MooWs& w = g_moo_ws[device & 63];
    double *dYT, *dF, *dR, *dC;
    int rc = upload_objective_major(w, h_Yc, n_cand, m, &dYT);
    if (rc != AOED_OK) return rc;
    std::vector<double> fs = sort_front(h_front, n_front, m);
    MCHK(ws_get(w, WS_FRONT, fs.size(), &dF));
    MCHK(ws_get(w, WS_REF, size_t(m), &dR));
    MCHK(ws_get(w, WS_CONTRIB, size_t(n_cand), &dC));
    if (!fs.empty()) MCU(cudaMemcpy(dF, fs.data(), fs.size() * sizeof(double), cudaMemcpyHostToDevice));
    MCU(cudaMemcpy(dR, h_ref, m * sizeof(double), cudaMemcpyHostToDevice));
    rc = launch_hvc(m, dYT, n_cand, dF, int(n_front), dR, nullptr, dC, nullptr);
    if (rc != AOED_OK) return rc;
    MCU(cudaMemcpy(h_contrib, dC, size_t(n_cand) * sizeof(double), cudaMemcpyDeviceToHost));
    return AOED_OK;
}

int aoed_hvi_select(int device, int64_t n_cand, int m, const double* h_Yc, int64_t n_front, const double* h_front,
                    const double* h_ref, int batch, const uint8_t* h_excluded, int64_t* h_idx, double* h_contrib) {
    if (n_cand < 0 || m < 1 || n_front < 0 || batch < 0 || !h_ref || !h_idx || (n_cand > 0 && !h_Yc) ||
        (n_front > 0 && !h_front))
        return mfail(AOED_ERR_INVALID, "bad argument");
    for (int b = 0; b < batch; ++b) {
        h_idx[b] = -1;
        if (h_contrib) h_contrib[b] = 0.0;
    }
    if (n_cand == 0 || batch == 0) return AOED_OK;
    if (m > 8) return mfail(AOED_ERR_UNSUPPORTED, "hypervolume improvement supports n_obj <= 8");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev || device < 0) return mfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    MCU(guard_.err);
    std::lock_guard<std::mutex> lock(g_moo_mu);
    MooWs& w = g_moo_ws[device & 63];
    const size_t front_cap = size_t(n_front + batch);
    const size_t smem = std::max<size_t>(front_cap, 1) * m * sizeof(double);
    if (smem > 200 * 1024) return mfail(AOED_ERR_UNSUPPORTED, "front + batch too large for the shared-memory front (n_obj * (n_front + batch) <= 25600)");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int grid = int(std::min<int64_t>((n_cand + HVI_THREADS - 1) / HVI_THREADS, int64_t(sms) * 8));

    double *dYT, *dF, *dR, *dC, *dOutV;
    uint8_t* dE;
    HviState* dSt;
    Top2* dPart;
    long long* dOutI;
    int rc = upload_objective_major(w, h_Yc, n_cand, m, &dYT);
    if (rc != AOED_OK) return rc;
    MCHK(ws_get(w, WS_FRONT, front_cap * m, &dF));
    MCHK(ws_get(w, WS_REF, size_t(m), &dR));
    MCHK(ws_get(w, WS_CONTRIB, size_t(n_cand), &dC));
    MCHK(ws_get(w, WS_EXCL, size_t(n_cand), &dE));
    MCHK(ws_get(w, WS_STATE, 1, &dSt));
    MCHK(ws_get(w, WS_PART, size_t(grid), &dPart));
    MCHK(ws_get(w, WS_OUT, size_t(4 * batch), &dOutV));  // [v1 | v2] doubles then [i1 | i2] 64-bit integers
    dOutI = reinterpret_cast<long long*>(dOutV + 2 * batch);
    MCU(cudaMemcpy(dR, h_ref, m * sizeof(double), cudaMemcpyHostToDevice));
    if (h_excluded) MCU(cudaMemcpy(dE, h_excluded, size_t(n_cand), cudaMemcpyHostToDevice));
    else MCU(cudaMemset(dE, 0, size_t(n_cand)));

    // host mirror of the greedy state: front in the reference's append order (hvi.py:42), taken candidates
    std::vector<double> front(h_front, h_front + size_t(n_front) * m);
    std::vector<uint8_t> taken(size_t(n_cand), 0);
    if (h_excluded) taken.assign(h_excluded, h_excluded + n_cand);
    // running hypervolume of the front: HV(front_0) once, then every accepted winner adds exactly its contribution — the scale
    // against which a gap between the two best candidates counts as a near-tie
    double hv_run = host_hypervolume(front, nullptr, m, h_ref);
    constexpr double kTieRel = 1e-12;
    static const bool no_guard = getenv("AOED_HVI_TIE_GUARD") && getenv("AOED_HVI_TIE_GUARD")[0] == '0';
    std::vector<double> v1(batch), v2(batch), contrib;
    std::vector<long long> i1(batch), i2(batch);

    int done = 0;  // rounds decided so far
    while (done < batch) {
        // ---- (re)start the device loop from the host state and run all remaining rounds without a host sync ----
        std::vector<double> fs = sort_front(front.data(), int64_t(front.size() / m), m);
        if (!fs.empty()) MCU(cudaMemcpy(dF, fs.data(), fs.size() * sizeof(double), cudaMemcpyHostToDevice));
        HviState st{int(front.size() / m), 0, 0, 0u};
        MCU(cudaMemcpy(dSt, &st, sizeof(st), cudaMemcpyHostToDevice));
        const int todo = batch - done;
        HviArgs a;
        a.YcT = dYT; a.n = n_cand; a.front = dF; a.ref = dR; a.excluded = dE; a.st = dSt; a.partial = dPart;
        a.out_v1 = dOutV; a.out_v2 = dOutV + batch; a.out_i1 = dOutI; a.out_i2 = dOutI + batch;
        for (int r = 0; r < todo; ++r) {
            rc = launch_hvi_round(m, a, grid, smem, nullptr);
            if (rc != AOED_OK) return rc;
        }
        MCU(cudaMemcpy(v1.data(), a.out_v1, todo * sizeof(double), cudaMemcpyDeviceToHost));
        MCU(cudaMemcpy(v2.data(), a.out_v2, todo * sizeof(double), cudaMemcpyDeviceToHost));
        MCU(cudaMemcpy(i1.data(), a.out_i1, todo * sizeof(long long), cudaMemcpyDeviceToHost));
        MCU(cudaMemcpy(i2.data(), a.out_i2, todo * sizeof(long long), cudaMemcpyDeviceToHost));
        MCU(cudaMemcpy(&st, dSt, sizeof(st), cudaMemcpyDeviceToHost));
        // ---- validate round by round: a near-tie between the two best is re-decided in the reference's form ----
        bool restart = false;
        for (int r = 0; r < st.round && !restart; ++r) {
            if (i1[r] < 0) {
                // No positive exclusive volume.  In the reference's form a candidate that no front point weakly dominates
                // can still come out positive by rounding of the two hypervolumes (e.g. one that defines the reference
                // point: zero-volume box, but it survives pymoo's non-dominated pre-filter and perturbs the summation);
                // weakly dominated ones give exactly zero.  Evaluate those few in that form before giving up.
                const double hv0 = host_hypervolume(front, nullptr, m, h_ref);
                const size_t kf = front.size() / m;
                double best = 0.0;
                long long best_i = -1;
                for (int64_t c = 0; c < n_cand; ++c) {
                    if (taken[c]) continue;
                    const double* y = h_Yc + c * m;
                    bool covered = false;
                    for (size_t f = 0; f < kf && !covered; ++f) {
                        bool le = true;
                        for (int t = 0; t < m; ++t) le = le && (front[f * m + t] <= y[t]);
                        covered = le;
                    }
                    if (covered) continue;
                    const double d = host_hypervolume(front, y, m, h_ref) - hv0;
                    if (d > best) {
                        best = d;
                        best_i = c;
                    }
                }
                if (best_i < 0) return AOED_OK;  // the caller draws the random pick (hvi.py:38-39)
                hv_run += best;
                h_idx[done] = best_i;
                if (h_contrib) h_contrib[done] = best;
                taken[size_t(best_i)] = 1;
                for (int t = 0; t < m; ++t) front.push_back(h_Yc[best_i * m + t]);
                ++done;
                restart = true;
                break;
            }
            long long win = i1[r];
            double win_v = v1[r];
            const double gap = v1[r] - std::max(v2[r], 0.0);
            if (!no_guard && gap <= kTieRel * (hv_run + v1[r])) {
                const double hv0 = host_hypervolume(front, nullptr, m, h_ref);
                {
                    // every candidate within the tolerance of the best takes part; HV(P u {y}) - HV(P) with strict '>' in
                    // index order (hvi.py:27-37)
                    std::vector<double> fsr = sort_front(front.data(), int64_t(front.size() / m), m);
                    if (!fsr.empty()) MCU(cudaMemcpy(dF, fsr.data(), fsr.size() * sizeof(double), cudaMemcpyHostToDevice));
                    std::vector<uint8_t> ex(taken);
                    MCU(cudaMemcpy(dE, ex.data(), size_t(n_cand), cudaMemcpyHostToDevice));
                    rc = launch_hvc(m, dYT, n_cand, dF, int(front.size() / m), dR, dE, dC, nullptr);
                    if (rc != AOED_OK) return rc;
                    contrib.resize(size_t(n_cand));
                    MCU(cudaMemcpy(contrib.data(), dC, size_t(n_cand) * sizeof(double), cudaMemcpyDeviceToHost));
                    const double thr = v1[r] - kTieRel * (hv0 + v1[r]);
                    double best = 0.0;
                    long long best_i = -1;
                    for (int64_t c = 0; c < n_cand; ++c) {
                        if (taken[c] || !(contrib[c] >= thr) || !(contrib[c] > 0.0)) continue;
                        const double d = host_hypervolume(front, h_Yc + c * m, m, h_ref) - hv0;
                        if (d > best) {
                            best = d;
                            best_i = c;
                        }
                    }
                    // no positive difference (contributions below the rounding resolution of the hypervolume): the
                    // reference falls through to its random pick (hvi.py:38-39) — so does the caller
                    if (best_i < 0) return AOED_OK;
                    win = best_i;
                    win_v = best;
                    restart = true;  // the device state was overwritten for the re-decision (and may have followed another winner)
                }
            }
            hv_run += win_v;
            h_idx[done] = win;
            if (h_contrib) h_contrib[done] = win_v;
            taken[size_t(win)] = 1;
            for (int t = 0; t < m; ++t) front.push_back(h_Yc[win * m + t]);
            ++done;
        }
        if (!restart) {
            if (st.stopped || st.round < todo) return AOED_OK;  // parked: the remaining entries stay -1
            break;
        }
        if (done < batch) MCU(cudaMemcpy(dE, taken.data(), size_t(n_cand), cudaMemcpyHostToDevice));
    }
    return AOED_OK;
}

}  // extern "C"
