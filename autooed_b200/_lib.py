"""ctypes binding of the C ABI in include/aoed_b200.h (libaoed_b200.so).

There is NO CPU fallback: importing this module without the built library, or using it without a CUDA
device, raises.  Build with `python -m autooed_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AOED_LIB") or os.path.join(_HERE, "lib", "libaoed_b200.so")  # AOED_LIB: tuning experiments

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOT_PD, ERR_UNSUPPORTED, ERR_NO_DEVICE = -1, -2, -3, -4, -5
EVAL_STD, EVAL_GRAD, EVAL_HESS, EVAL_EXACT = 1, 2, 4, 8
ACQ_NONE, ACQ_IDENTITY, ACQ_UCB, ACQ_EI, ACQ_PI = 0, 1, 2, 3, 4

_c_double_p = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p

# every exported symbol of include/aoed_b200.h: name -> (restype, argtypes)
SIGNATURES = {
    "aoed_abi_version": (ctypes.c_int, []),
    "aoed_last_error": (ctypes.c_char_p, []),
    "aoed_device_count": (ctypes.c_int, []),
    "aoed_host_alloc": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_uint64]),
    "aoed_host_free": (ctypes.c_int, [_vp]),
    "aoed_gp_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "aoed_gp_destroy": (ctypes.c_int, [_vp]),
    "aoed_gp_set_train": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "aoed_gp_lml_batch": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "aoed_gp_set_theta": (ctypes.c_int, [_vp, _vp]),
    "aoed_gp_get_state": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp]),
    "aoed_gp_eval": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_uint32, ctypes.c_int, _vp] + [_vp] * 9),
    "aoed_gp_eval_device": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_uint32, ctypes.c_int, _vp] + [_vp] * 9 + [_vp]),
    "aoed_gp_set_profiling": (ctypes.c_int, [_vp, ctypes.c_int]),
    "aoed_gp_get_counters": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int64), _c_double_p,
                                            ctypes.POINTER(ctypes.c_int64), _c_double_p]),
    "aoed_gp_reset_counters": (ctypes.c_int, [_vp]),
    "aoed_gp_get_fit_counters": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "aoed_nsga2_run": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int64, _vp, _vp, _vp,
                       ctypes.c_int, ctypes.c_int, ctypes.c_uint64, _vp, _vp, ctypes.POINTER(ctypes.c_int64),
                       ctypes.POINTER(ctypes.c_int64)]),
    "aoed_nsga2_run_rff": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [_vp] * 6 +
                           [ctypes.c_int64, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, _vp, _vp,
                            ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "aoed_nsga2_survival": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, _vp, ctypes.c_int64, _vp, _vp, _vp]),
    "aoed_rff_eval": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [_vp] * 7 +
                      [ctypes.c_uint32, _vp, _vp, _vp]),
    "aoed_pareto_mask": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, _vp, _vp]),
    "aoed_pareto_mask_rows": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, _vp, ctypes.c_int64, ctypes.c_int64, _vp]),
    "aoed_pareto_rank": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, _vp, _vp]),
    "aoed_hvi_select": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, _vp, ctypes.c_int64, _vp, _vp,
                                       ctypes.c_int, _vp, _vp, _vp]),
    "aoed_hv_contributions": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int, _vp, ctypes.c_int64, _vp, _vp, _vp]),
    "aoed_graph_cut": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp,
                                      ctypes.POINTER(ctypes.c_int64)]),
}

_lib = None


class AoedError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libaoed_b200 error {code}: {msg}")
        self.code = code


def load():
    """Loads the shared library (once).  Raises if it has not been built — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA extension with `python -m autooed_b200.build` "
            "(autooed_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().aoed_last_error()
        msg = msg.decode() if msg else ""
        if rc == ERR_NOT_PD:
            raise np.linalg.LinAlgError(msg)  # what sklearn raises (_gpr.py:352-361)
        if rc == ERR_INVALID:
            raise AssertionError(msg)  # the reference's assert on unfitted models / bad arguments
        raise AoedError(rc, msg)


def device_count():
    return load().aoed_device_count()


def require_device():
    n = device_count()
    if n <= 0:
        raise AoedError(ERR_NO_DEVICE, "no CUDA device visible; autooed_b200 has no CPU fallback")
    return n


def ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_PINNED_MIN_BYTES = 1 << 20
_POOL_CAP_BYTES = 4 << 30
_pool = {}          # nbytes -> [pointer, ...] of free pinned blocks (cudaHostAlloc costs ~0.3 ms/MB: never pay it twice)
_pool_bytes = 0


def _release(lib, p, nbytes):
    global _pool_bytes
    if _pool_bytes + nbytes <= _POOL_CAP_BYTES:
        _pool.setdefault(nbytes, []).append(p)
        _pool_bytes += nbytes
    else:
        lib.aoed_host_free(p)


def empty(shape, dtype=np.float64):
    """Result array; large ones are backed by (pooled) pinned host memory so the D2H copy is a true async DMA."""
    global _pool_bytes
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if nbytes < _PINNED_MIN_BYTES:
        return np.empty(shape, dtype=dtype)
    lib = load()
    free = _pool.get(nbytes)
    if free:
        addr = free.pop()
        _pool_bytes -= nbytes
    else:
        p = _vp()
        check(lib.aoed_host_alloc(ctypes.byref(p), nbytes))
        addr = p.value
    buf = (ctypes.c_char * nbytes).from_address(addr)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    weakref.finalize(buf, _release, lib, addr, nbytes)
    return arr
