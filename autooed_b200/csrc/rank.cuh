// One-CTA non-dominated sort shared by the Pareto C ABI (moo.cu) and the device-resident NSGA-II (nsga2.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

namespace aoed {

// Whole non-dominated sort in ONE CTA for populations that fit in shared memory (the NSGA-II regime, n <= 4096):
// domination counts, then front peeling with the counts decremented by the members of each peeled front — no host
// round trip per front.  Same semantics as the multi-launch path below (rank = front number, pymoo NonDominatedSorting).
constexpr int RS_THREADS = 1024;

// slot = (*ctr)++ for the lanes with `pred`, one shared-memory atomic per warp instead of one per lane (a thousand
// same-address atomics serialise).  Must be reached by all 32 lanes; returns -1 for lanes without `pred`.
__device__ __forceinline__ int warp_slot(int* ctr, bool pred) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return -1;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}
constexpr int RS_MAX_N = 4096;

// domination counts of all n points over many CTAs: a warp owns 32 consecutive points; the candidates for "dominates
// it" are cut into DC_SPLIT contiguous slices (grid.y), staged in shared memory — read straight from global memory every
// step of the scan is a fresh cache line and pays an L2 round trip — and scanned by the DC_WARPS warps of the CTA
// (every lane reads the same j: broadcast).  cnt[y * n + i] = partial count of slice y; the sort kernel adds the slices.
constexpr int DC_WARPS = 8;
constexpr int DC_SPLIT = 4;
inline int dom_count_chunk(int n) { return (n + DC_SPLIT - 1) / DC_SPLIT; }

template <int M>
__global__ void __launch_bounds__(32 * DC_WARPS) dom_count_kernel(const double* __restrict__ YT, int n, int chunk, int* __restrict__ cnt) {
    extern __shared__ double dc_smem[];  // [M][chunk]
    __shared__ int part[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = blockIdx.x * 32 + lane;
    const int j0 = blockIdx.y * chunk, len = min(chunk, n - j0);
    for (int t = threadIdx.x; t < M * chunk; t += 32 * DC_WARPS) {
        const int k = t / chunk, j = t - k * chunk;
        if (j < len) dc_smem[t] = YT[int64_t(k) * n + j0 + j];
    }
    if (w == 0) part[lane] = 0;
    __syncthreads();
    if (i < n) {
        double yi[M];
#pragma unroll
        for (int k = 0; k < M; ++k) yi[k] = YT[int64_t(k) * n + i];
        int c = 0;
#pragma unroll 4
        for (int j = w; j < len; j += DC_WARPS) {
            bool le = true, lt = false;
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const double v = dc_smem[k * chunk + j];
                le = le && (v <= yi[k]);
                lt = lt || (v < yi[k]);
            }
            c += (le && lt) ? 1 : 0;
        }
        atomicAdd(&part[lane], c);
    }
    __syncthreads();
    if (w == 0 && i < n) cnt[int64_t(blockIdx.y) * n + i] = part[lane];
}

// `cnt_in` (optional): domination counts from dom_count_kernel — worth the extra launch for n above ~1000, where the
// single CTA would spend most of its time on the n^2 counting.  `need`: stop once that many points are ranked (NSGA-II
// survival only needs the fronts that fill the next population, pymoo `n_stop_if_ranked`); the rest get the next rank
// number.  need >= n gives the full sort.  `last_rank_out` (optional) receives the last rank that is a real front.
template <int M>
__global__ void __launch_bounds__(RS_THREADS) rank_small_kernel(const double* __restrict__ YT, int n, int32_t* __restrict__ rank_out,
                                                                const int* __restrict__ cnt_in, int need,
                                                                int* __restrict__ last_rank_out) {
    extern __shared__ double rs_smem[];
    double* Y = rs_smem;                                   // [M][n]
    int* cnt = reinterpret_cast<int*>(Y + size_t(M) * n);  // [n] number of not-yet-peeled points dominating i
    int* rk = cnt + n;                                     // [n]
    int* front = rk + n;                                   // [n] compacted current front
    __shared__ int n_front, n_left, last_full;
    const int tid = threadIdx.x;
    for (int i = tid; i < M * n; i += RS_THREADS) Y[i] = YT[i];
    if (tid == 0) {
        n_left = n;
        last_full = 0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += RS_THREADS) {
        int c = 0;
        if (cnt_in) {
#pragma unroll
            for (int y = 0; y < DC_SPLIT; ++y) c += cnt_in[y * n + i];
        } else {
            double yi[M];
#pragma unroll
            for (int k = 0; k < M; ++k) yi[k] = Y[k * n + i];
            for (int j = 0; j < n; ++j) {
                bool le = true, lt = false;
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const double v = Y[k * n + j];
                    le = le && (v <= yi[k]);
                    lt = lt || (v < yi[k]);
                }
                c += (le && lt) ? 1 : 0;
            }
        }
        cnt[i] = c;
        rk[i] = -1;
    }
    __syncthreads();
    for (int r = 0; n_left > 0; ++r) {
        if (tid == 0) n_front = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += RS_THREADS) {
            const int i = i0 + tid;
            const bool in_front = i < n && rk[i] < 0 && cnt[i] == 0;
            const int slot = warp_slot(&n_front, in_front);
            if (in_front) {
                rk[i] = r;
                front[slot] = i;
            }
        }
        __syncthreads();
        const int nf = n_front;
        if (nf == 0) break;  // cannot happen for finite inputs (a minimal element always exists); NaN rows stay at -1
        if (n - n_left + nf >= need && nf < n_left) {  // enough ranked: everything left shares the next rank number
            for (int i = tid; i < n; i += RS_THREADS)
                if (rk[i] < 0) rk[i] = r + 1;
            if (tid == 0) last_full = r;
            __syncthreads();
            break;
        }
        if (tid == 0) last_full = r;
        for (int q = tid; q < n; q += RS_THREADS) {
            if (rk[q] >= 0) continue;
            double yq[M];
#pragma unroll
            for (int k = 0; k < M; ++k) yq[k] = Y[k * n + q];
            int dec = 0;
            for (int t = 0; t < nf; ++t) {
                const int pidx = front[t];
                bool le = true, lt = false;
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const double v = Y[k * n + pidx];
                    le = le && (v <= yq[k]);
                    lt = lt || (v < yq[k]);
                }
                dec += (le && lt) ? 1 : 0;
            }
            cnt[q] -= dec;
        }
        __syncthreads();
        if (tid == 0) n_left -= nf;
        __syncthreads();
    }
    for (int i = tid; i < n; i += RS_THREADS) rank_out[i] = rk[i];
    if (last_rank_out && tid == 0) *last_rank_out = last_full;  // highest rank whose front was peeled completely
}


inline size_t rank_small_smem(int m, int n) { return size_t(m) * n * sizeof(double) + 3 * size_t(n) * sizeof(int); }
constexpr int RS_SPLIT_N = 1024;  // above this the domination counts come from dom_count_kernel

// launches the sort of n <= RS_MAX_N points on stream s; `cnt_scratch` ([DC_SPLIT * n] ints on the device) enables the split
// counting for large n (pass nullptr to stay with the single kernel)
constexpr size_t RS_SMEM_LIMIT = 200 * 1024;  // callers check rank_small_smem(m, n) against this

template <int M>
cudaError_t launch_rank_small_m(const double* YT, int n, int32_t* rank, int* cnt_scratch, int need, cudaStream_t s,
                                int* last_rank_out = nullptr) {
    const size_t sm = rank_small_smem(M, n);
    const int chunk = dom_count_chunk(n);
    const size_t dsm = size_t(M) * chunk * sizeof(double);
    static thread_local int opted_in_device = -1;  // the shared-memory opt-in is per device and costs a driver call
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (opted_in_device != dev) {
        e = cudaFuncSetAttribute(rank_small_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(RS_SMEM_LIMIT));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(dom_count_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(size_t(M) * dom_count_chunk(RS_MAX_N) * sizeof(double)));
        if (e != cudaSuccess) return e;
        opted_in_device = dev;
    }
    const int* cnt_in = nullptr;
    if (cnt_scratch && n > RS_SPLIT_N) {
        dom_count_kernel<M><<<dim3((n + 31) / 32, DC_SPLIT), 32 * DC_WARPS, dsm, s>>>(YT, n, chunk, cnt_scratch);
        cnt_in = cnt_scratch;
    }
    rank_small_kernel<M><<<1, RS_THREADS, sm, s>>>(YT, n, rank, cnt_in, need, last_rank_out);
    return cudaGetLastError();
}

}  // namespace aoed
