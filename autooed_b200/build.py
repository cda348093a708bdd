"""Builds libaoed_b200.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

    python -m autooed_b200.build            # or __graft_entry__.build()

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libaoed_b200.so")
SOURCES = ["gp.cu", "launch_xcov.cu", "launch_fitk.cu", "launch_small_eval.cu", "launch_hess.cu", "blocked.cu", "fit_launch.cu",
           "fit_launch_nu1.cu", "fit_launch_nu3.cu", "fit_launch_nu5.cu", "fit_launch_rbf.cu", "moo.cu", "rff.cu", "nsga2.cu", "graphcut.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "aoed_b200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--threads", "0"]  # the sources compile concurrently
    for macro in ("AOED_XCOV_UNROLL", "AOED_XCOV_MINBLOCKS", "AOED_XCOV_XS_SMEM"):  # tuning experiments; defaults live in posterior.cuh
        if os.environ.get(macro):
            cmd += [f"-D{macro}={os.environ[macro]}"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    out = os.environ.get("AOED_BUILD_OUT", LIB)  # tuning variants are built next to the library and selected with AOED_LIB
    cmd += ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libaoed_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
