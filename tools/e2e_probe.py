"""Where does the host-pointer API lose time against the device-resident API?  (tuning aid)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from autooed_b200 import GaussianProcess, UpperConfidenceBound, _lib
from autooed_b200.problem import SimpleProblem
N, d, m, C = 2000, 20, 3, 1 << 19
X, Y, thetas = bench.make_problem(N, d, m)
gp = GaussianProcess(SimpleProblem(d, m), nu=5)
gp.fit_fixed(X, Y, thetas)
acq = UpperConfidenceBound(gp); acq.fit(X, Y)
Xs = bench.candidates(C, d, 2)
dev = torch.device("cuda:0")
dX = torch.from_numpy(Xs).to(dev)
dA = torch.empty((C, m), dtype=torch.float64, device=dev); ddA = torch.empty((C, m, d), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

t_dev = timeit(lambda: gp.evaluate_device(dX, C, std=True, gradient=True, acq_kind=_lib.ACQ_UCB, acq_param=[float(N)], d_A=dA, d_dA=ddA, stream=st))
t_full = timeit(lambda: acq.evaluate(Xs, dtype='normalized', gradient=True))
# value + gradient computed, but only the value copied back: isolates the D2H of dA
import ctypes
lib, h = _lib.load(), gp._ensure_handle()
A = _lib.empty((C, m)); param = np.array([float(N)])
def only_A():
    _lib.check(lib.aoed_gp_eval(h, C, _lib.ptr(Xs), _lib.EVAL_STD | _lib.EVAL_GRAD, _lib.ACQ_UCB, _lib.ptr(param),
                                None, None, None, None, None, None, _lib.ptr(A), None, None))
t_A = timeit(only_A)
pinX = _lib.empty((C, d)); pinX[:] = Xs
t_pin = timeit(lambda: acq.evaluate(pinX, dtype='normalized', gradient=True))
print(f"chunk={os.environ.get('AOED_CHUNK','default')} device {t_dev:.1f} ms | host full {t_full:.1f} | host, A only {t_A:.1f} | host, pinned X {t_pin:.1f}")
