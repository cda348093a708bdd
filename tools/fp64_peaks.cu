// FP64 pipe probe for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) issue-bound throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peaks tools/fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: 256 FMA per warp instruction
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8: 1024 FMA per warp instruction (A: 4 regs, B: 2 regs, C: 4 regs)
__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16: 2048 FMA per warp instruction (A: 8 regs, B: 4 regs, C: 4 regs)
__global__ void dmma16816_kernel(double* out, int iters, double a, double b) {
    double c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// exp() throughput in FP64 (library exp) per thread
__global__ void dexp_kernel(double* out, int iters, double a) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = -1e-3 * (threadIdx.x + i);
    double s = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { s += exp(x[i]); x[i] -= a; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 4096;
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int threads = 256; int blocks = sms * warps_per_sm * 32 / threads;
        float ms;
        ms = timeit([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.000001, 1e-9); }, 5);
        printf(", \"dfma_w%d_tflops\": %.2f", warps_per_sm, 2.0 * 16 * iters * blocks * threads / ms * 1e-9);
        ms = timeit([&] { dmma884_kernel<<<blocks, threads>>>(out, iters, 1.000001, 1e-9); }, 5);
        printf(", \"dmma884_w%d_tflops\": %.2f", warps_per_sm, 2.0 * 256 * 8 * iters * (blocks * threads / 32.0) / ms * 1e-9);
        ms = timeit([&] { dmma1688_kernel<<<blocks, threads>>>(out, iters, 1.000001, 1e-9); }, 5);
        printf(", \"dmma1688_w%d_tflops\": %.2f", warps_per_sm, 2.0 * 1024 * 8 * iters * (blocks * threads / 32.0) / ms * 1e-9);
        ms = timeit([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters, 1.000001, 1e-9); }, 5);
        printf(", \"dmma16816_w%d_tflops\": %.2f", warps_per_sm, 2.0 * 2048 * 8 * iters * (blocks * threads / 32.0) / ms * 1e-9);
    }
    {
        int threads = 256, blocks = sms * 8;
        float ms = timeit([&] { dexp_kernel<<<blocks, threads>>>(out, 1024, 1e-4); }, 5);
        printf(", \"dexp_gexp_per_s\": %.2f", 4.0 * 1024 * blocks * threads / ms * 1e-6);
    }
    CK(cudaDeviceSynchronize());
    printf("}\n");
    return 0;
}
