// One-CTA non-dominated sort shared by the Pareto C ABI (moo.cu) and the device-resident NSGA-II (nsga2.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aoed {

// Whole non-dominated sort in ONE CTA for populations that fit in shared memory (the NSGA-II regime, n <= 4096):
// domination counts, then front peeling with the counts decremented by the members of each peeled front — no host
// round trip per front.  Same semantics as the multi-launch path below (rank = front number, pymoo NonDominatedSorting).
constexpr int RS_THREADS = 1024;
constexpr int RS_MAX_N = 4096;

template <int M>
__global__ void __launch_bounds__(RS_THREADS) rank_small_kernel(const double* __restrict__ YT, int n, int32_t* __restrict__ rank_out) {
    extern __shared__ double rs_smem[];
    double* Y = rs_smem;                                   // [M][n]
    int* cnt = reinterpret_cast<int*>(Y + size_t(M) * n);  // [n] number of not-yet-peeled points dominating i
    int* rk = cnt + n;                                     // [n]
    int* front = rk + n;                                   // [n] compacted current front
    __shared__ int n_front, n_left;
    const int tid = threadIdx.x;
    for (int i = tid; i < M * n; i += RS_THREADS) Y[i] = YT[i];
    if (tid == 0) n_left = n;
    __syncthreads();
    for (int i = tid; i < n; i += RS_THREADS) {
        double yi[M];
#pragma unroll
        for (int k = 0; k < M; ++k) yi[k] = Y[k * n + i];
        int c = 0;
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
        cnt[i] = c;
        rk[i] = -1;
    }
    __syncthreads();
    for (int r = 0; n_left > 0; ++r) {
        if (tid == 0) n_front = 0;
        __syncthreads();
        for (int i = tid; i < n; i += RS_THREADS)
            if (rk[i] < 0 && cnt[i] == 0) {
                rk[i] = r;
                front[atomicAdd(&n_front, 1)] = i;
            }
        __syncthreads();
        const int nf = n_front;
        if (nf == 0) break;  // cannot happen for finite inputs (a minimal element always exists); NaN rows stay at -1
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
}


inline size_t rank_small_smem(int m, int n) { return size_t(m) * n * sizeof(double) + 3 * size_t(n) * sizeof(int); }

}  // namespace aoed
