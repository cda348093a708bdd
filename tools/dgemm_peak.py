"""cuBLAS FP64 GEMM peak on this GPU (roofline denominator for the FP64 tensor/FMA pipe).
Prints one JSON line; run under gpurun."""
import json, sys, time
import torch

def main():
    dev = torch.device("cuda:0")
    out = {"gpu": torch.cuda.get_device_name(0)}
    for n in (4096, 8192):
        a = torch.randn(n, n, device=dev, dtype=torch.float64)
        b = torch.randn(n, n, device=dev, dtype=torch.float64)
        c = torch.empty_like(a)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"dgemm_{n}_tflops_burst"] = 2.0 * n ** 3 / best * 1e-9
        # sustained ~3 s
        t0 = time.time(); reps = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 3.0:
            for _ in range(4):
                torch.matmul(a, b, out=c)
            reps += 4
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        out[f"dgemm_{n}_tflops_sustained"] = 2.0 * n ** 3 * reps / e0.elapsed_time(e1) * 1e-9
    # tall-skinny shape of the posterior TRMM: (C x N) x (N x N), N = 2000
    C, N = 16384, 2000
    a = torch.randn(C, N, device=dev, dtype=torch.float64)
    b = torch.randn(N, N, device=dev, dtype=torch.float64)
    c = torch.empty(C, N, device=dev, dtype=torch.float64)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["dgemm_16384x2000x2000_tflops"] = 2.0 * C * N * N / best * 1e-9
    print(json.dumps(out))

if __name__ == "__main__":
    main()
