// Times cusolverDnDpotrf vs cusolverDnDpotrfBatched and cublasDtrsm vs cublasDtrsmBatched at mid sizes (tuning aid for the
// multi-stream LML path).  nvcc -O2 -arch=sm_100a potrf_probe.cu -lcusolver -lcublas -o potrf_probe
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdio>
#include <vector>

int main() {
    cusolverDnHandle_t sol; cublasHandle_t blas;
    cusolverDnCreate(&sol); cublasCreate(&blas);
    const int B = 2;
    for (int N : {256, 300, 500, 1000}) {
        std::vector<double> h(size_t(N) * N);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) h[size_t(i) * N + j] = (i == j ? N : 0.0) + 1.0 / (1.0 + abs(i - j));
        double *dA[B], *dI[B], **dAp, **dIp; int* info;
        for (int b = 0; b < B; ++b) { cudaMalloc(&dA[b], h.size() * 8); cudaMalloc(&dI[b], h.size() * 8); }
        cudaMalloc(&dAp, B * sizeof(double*)); cudaMalloc(&dIp, B * sizeof(double*)); cudaMalloc(&info, B * sizeof(int));
        cudaMemcpy(dAp, dA, B * sizeof(double*), cudaMemcpyHostToDevice); cudaMemcpy(dIp, dI, B * sizeof(double*), cudaMemcpyHostToDevice);
        int lwork = 0; cusolverDnDpotrf_bufferSize(sol, CUBLAS_FILL_MODE_UPPER, N, dA[0], N, &lwork);
        double* work; cudaMalloc(&work, size_t(lwork) * 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float t_single = 0, t_batched = 0, t_trsm = 0, t_trsmb = 0;
        const double one = 1.0;
        for (int rep = 0; rep < 6; ++rep) {
            for (int b = 0; b < B; ++b) cudaMemcpy(dA[b], h.data(), h.size() * 8, cudaMemcpyHostToDevice);
            cudaEventRecord(e0);
            for (int b = 0; b < B; ++b) cusolverDnDpotrf(sol, CUBLAS_FILL_MODE_UPPER, N, dA[b], N, work, lwork, info);
            cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) t_single += ms;
            for (int b = 0; b < B; ++b) cudaMemset(dI[b], 0, h.size() * 8);
            cudaEventRecord(e0);
            for (int b = 0; b < B; ++b) cublasDtrsm(blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, N, N, &one, dA[b], N, dI[b], N);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (rep) t_trsm += ms;
            cudaEventRecord(e0);
            cublasDtrsmBatched(blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, N, N, &one, (const double* const*)dAp, N, dIp, N, B);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (rep) t_trsmb += ms;
            for (int b = 0; b < B; ++b) cudaMemcpy(dA[b], h.data(), h.size() * 8, cudaMemcpyHostToDevice);
            cudaEventRecord(e0);
            cusolverDnDpotrfBatched(sol, CUBLAS_FILL_MODE_UPPER, N, dAp, N, info, B);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (rep) t_batched += ms;
        }
        printf("N %4d batch %d: potrf x%d %.3f ms | potrfBatched %.3f ms | trsm x%d %.3f ms | trsmBatched %.3f ms\n", N, B, B, t_single / 5,
               t_batched / 5, B, t_trsm / 5, t_trsmb / 5);
    }
    return 0;
}
