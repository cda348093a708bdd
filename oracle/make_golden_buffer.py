"""TEST INFRASTRUCTURE ONLY — pins the ParetoDiscovery shell (performance buffer insert / move_origin /
_update_cell / sample / flattened, stochastic sampling, family-aware batch proposal) to the reference: runs the UNMODIFIED
`/root/reference/autooed/mobo/solver/pareto_discovery/buffer.py:10-183,306-398`, `pareto_discovery.py:529-560` and
`utils.py:5-134` on seeded inputs and writes `tests/golden/pdshell_cases.npz`.

pygco / pymoo are absent: `pygco.cut_from_graph` is only reached by `sparse_approximation` (not recorded), and
`pymoo.factory.get_performance_indicator('hv', ref_point=r)` is stood in by the oracle's exact hypervolume (so the LOOP
logic of the proposal functions — family cycles, masks, padding, random fallback — is pinned to the reference, the
hypervolume arithmetic itself is the oracle's, see oracle/moo_oracle.py).

Run in the authoring container only:   python oracle/make_golden_buffer.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import moo_oracle as mo  # noqa: E402
from oracle.make_golden_discovery import load_reference_module  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class _HV:
    def __init__(self, ref_point):
        self.ref = np.asarray(ref_point, dtype=float)

    def calc(self, F):
        return mo.hypervolume(F, self.ref)


def dump_cells(rec, key, buf):
    """Cell contents as flat arrays + offsets."""
    sizes = np.array([len(c) for c in buf.buffer_dist])
    rec[f"{key}_sizes"] = sizes
    rec[f"{key}_origin"] = np.array(buf.origin)
    rec[f"{key}_count"] = np.int64(buf.sample_count)
    if sizes.sum() == 0:
        return
    rec[f"{key}_x"] = np.vstack([np.array(c) for c in buf.buffer_x if len(c)])
    rec[f"{key}_y"] = np.vstack([np.array(c) for c in buf.buffer_y if len(c)])
    rec[f"{key}_dist"] = np.concatenate([np.array(c) for c in buf.buffer_dist if len(c)])
    rec[f"{key}_pid"] = np.concatenate([np.array(c) for c in buf.buffer_patch_id if len(c)]).astype(np.int64)


def main():
    _, pd = load_reference_module()
    import importlib
    sys.modules['pymoo.factory'].get_performance_indicator = lambda name, ref_point: _HV(ref_point)
    buf_mod = importlib.import_module('autooed.mobo.solver.pareto_discovery.buffer')
    utils = importlib.import_module('autooed.mobo.solver.pareto_discovery.utils')
    utils.get_performance_indicator = lambda name, ref_point: _HV(ref_point)
    rec = {}
    for tag, m, cell_num, cell_size in (("b2", 2, 20, 3), ("b3", 3, 30, 2), ("b2u", 2, 12, None)):
        rng = np.random.default_rng(len(tag) * 7 + m)
        np.random.seed(17)
        buf = buf_mod.get_buffer(m, cell_num, cell_size=cell_size, origin=None, origin_constant=1e-2, delta_b=0.2, label_cost=10)
        d = 4
        rec[f"{tag}_cfg"] = np.array([m, cell_num, -1 if cell_size is None else cell_size, d])
        n_ins = 5
        for k in range(n_ins):
            n = 40
            X = rng.random((n, d))
            Y = rng.random((n, m)) * (1.0 + 0.3 * k) + 0.05
            if k == 1:
                Y[3] = -0.05            # below the origin: the buffer moves its origin and re-bins everything
            if k == 3:
                Y[7, 0] = buf.origin[0]  # touches the origin in one coordinate: moves as well (:149)
                Y[:5] = Y[5:10]          # ties in distance inside a cell
            pid = rng.integers(0, 6, n) + 10 * k
            if k == 0:
                pid = pid[:25]           # the reference's first insert passes fewer patch ids than rows (:460-462)
            rec[f"{tag}_in{k}_X"], rec[f"{tag}_in{k}_Y"], rec[f"{tag}_in{k}_pid"] = X, Y, pid
            buf.insert(X, Y, pid)
            dump_cells(rec, f"{tag}_after{k}", buf)
        rec[f"{tag}_n_ins"] = np.int64(n_ins)
        for n in (5, 18, 60, 400):
            np.random.seed(100 + n)
            rec[f"{tag}_sample{n}"] = buf.sample(n)
        fx, fy = buf.flattened()
        rec[f"{tag}_flat_x"], rec[f"{tag}_flat_y"] = fx, fy
        if m == 3:
            rec[f"{tag}_cell_vecs"] = buf.cell_vecs

    # stochastic sampling (pareto_discovery.py:529-560) on a fake algorithm object
    class _Pop:
        def __init__(self, X):
            self.X = X

        def get(self, key):
            return self.X

    rng = np.random.default_rng(5)
    xs = rng.random((9, 5))
    for tag, constr in (("ss_free", None), ("ss_constr", lambda X: X[:, 0] + X[:, 1] - 1.2)):
        fake = types.SimpleNamespace(pop=_Pop(xs), delta_p=10, constr_func=constr,
                                     problem=types.SimpleNamespace(xl=np.zeros(5), xu=np.ones(5)))
        np.random.seed(23)
        rec[f"{tag}_out"] = pd.ParetoDiscoveryAlgorithm._stochastic_sampling(fake)
    rec["ss_xs"] = xs

    # batch proposal (utils.py:5-134)
    for tag, m, n_c, batch in (("pnb2", 2, 14, 6), ("pnb3", 3, 20, 9), ("pnb_small", 2, 4, 7)):
        rng = np.random.default_rng(m * 11 + n_c)
        front = mo.find_pareto_front(rng.random((30, m)) * 0.6 + 0.4)
        pf = rng.random((n_c, m)) * 0.9
        pf[-3:] = pf[:3]                     # duplicates: exact ties, lowest index wins
        ps = rng.random((n_c, 3))
        labels = rng.integers(0, 4, n_c)
        ref = np.max(np.vstack([front, pf]), axis=0)
        np.random.seed(3)
        Xn, Yn, fam = utils.propose_next_batch(front, ref, pf, ps, batch, labels)
        np.random.seed(3)
        Xn2, Yn2 = utils.propose_next_batch_without_label(front, ref, pf, ps, batch)
        rec.update({f"{tag}_front": front, f"{tag}_pf": pf, f"{tag}_ps": ps, f"{tag}_labels": labels, f"{tag}_ref": ref,
                    f"{tag}_batch": np.int64(batch), f"{tag}_X": Xn, f"{tag}_Y": Yn, f"{tag}_fam": np.array(fam),
                    f"{tag}_X_nolabel": Xn2, f"{tag}_Y_nolabel": Yn2})
    np.savez_compressed(os.path.join(OUT, "pdshell_cases.npz"), **rec)
    print("wrote pdshell_cases.npz with", len(rec), "arrays")


if __name__ == "__main__":
    main()
