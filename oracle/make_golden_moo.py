"""TEST INFRASTRUCTURE ONLY — pins the Pareto helpers to the reference: executes the UNMODIFIED
`/root/reference/autooed/utils/pareto.py` (`find_pareto_front` :27-46, `check_pareto` :49-62) on seeded inputs and writes
`tests/golden/pareto_cases.npz`.  That module imports `pymoo.performance_indicator.hv.Hypervolume` at the top (pareto.py:7)
for `calc_hypervolume` only; pymoo is not installable here, so an empty stub of that one third-party module is registered
in `sys.modules` first — the two functions exercised here never touch it.

Run in the authoring container only:   python oracle/make_golden_moo.py
"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "pareto_cases.npz")


def load_reference_pareto():
    for name in ("pymoo", "pymoo.performance_indicator", "pymoo.performance_indicator.hv"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pymoo.performance_indicator.hv"].Hypervolume = object
    rl._stub("autooed", os.path.join(rl.REF_ROOT, "autooed"))
    rl._stub("autooed.utils", os.path.join(rl.REF_ROOT, "autooed", "utils"))
    if rl.REF_ROOT not in sys.path:
        sys.path.insert(0, rl.REF_ROOT)
    return importlib.import_module("autooed.utils.pareto")


def cases():
    rng = np.random.default_rng(42)
    out = {}
    out["rand_m2"] = rng.random((200, 2))
    out["rand_m3"] = rng.random((300, 3))
    out["rand_m4"] = rng.random((150, 4))
    out["grid_ties_m2"] = np.round(rng.random((120, 2)), 1)        # many equal coordinates and duplicate rows
    out["grid_ties_m3"] = np.round(rng.random((200, 3)), 1)
    out["dups_m3"] = np.repeat(rng.random((20, 3)), 3, axis=0)     # every point three times (nobody dominates a copy)
    out["single"] = rng.random((1, 3))
    out["line_m2"] = np.stack([np.linspace(0, 1, 50), 1 - np.linspace(0, 1, 50)], 1)  # all non-dominated
    out["chain_m2"] = np.stack([np.linspace(0, 1, 50), np.linspace(0, 1, 50)], 1)      # one point dominates all
    return out


def main():
    ref = load_reference_pareto()
    bundle = {}
    for name, Y in cases().items():
        front, idx = ref.find_pareto_front(Y, return_index=True)
        bundle[f"{name}__Y"] = Y
        bundle[f"{name}__front"] = np.asarray(front)
        bundle[f"{name}__idx"] = np.asarray(idx, dtype=np.int64)
        bundle[f"{name}__mask"] = np.asarray(ref.check_pareto(Y), dtype=bool)
    Y = cases()["rand_m3"]
    front, idx = ref.find_pareto_front(Y, return_index=True, obj_type=["min", "max", "min"])
    bundle["objtype__Y"] = Y
    bundle["objtype__front"] = np.asarray(front)
    bundle["objtype__idx"] = np.asarray(idx, dtype=np.int64)
    bundle["objtype__mask"] = np.asarray(ref.check_pareto(Y, obj_type=["min", "max", "min"]), dtype=bool)
    np.savez_compressed(OUT, **bundle)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
