"""Hypervolume-improvement selection — drop-in for `autooed/mobo/selection/hvi.py:16-47`.
The batch x |candidates| exact-hypervolume loop of the reference runs as device-resident greedy rounds (`aoed_hvi_select`:
one kernel per round, candidates / front / taken mask stay on the GPU, no host synchronisation between rounds); a near-tie
between the two best candidates of a round is re-decided inside the library with the reference's own
`HV(front + y) - HV(front)` form (hvi.py:28-37); only the random fallback (hvi.py:38-39) stays here because it consumes
the global numpy RNG."""
import numpy as np

from autooed_b200 import _lib
from autooed_b200.selection.base import Selection
from autooed_b200.utils.pareto import find_pareto_front


def hvi_select_indices(Y_candidate, Y, batch_size, device=0):
    Yc = _lib.as_f64(Y_candidate)
    Y = np.asarray(Y, dtype=np.float64)
    lib = _lib.load()
    _lib.require_device()
    m = Yc.shape[1]
    curr_pfront = find_pareto_front(Y, device=device).reshape(-1, m)
    ref_point = _lib.as_f64(np.max(np.vstack([Yc, Y]), axis=0))  # hvi.py:20
    excluded = np.zeros(len(Yc), dtype=np.uint8)
    chosen = []
    while len(chosen) < batch_size:
        want = batch_size - len(chosen)
        idx = np.full(want, -1, dtype=np.int64)
        front = _lib.as_f64(curr_pfront)
        _lib.check(lib.aoed_hvi_select(device, Yc.shape[0], m, _lib.ptr(Yc), front.shape[0], _lib.ptr(front),
                                       _lib.ptr(ref_point), want, _lib.ptr(excluded), _lib.ptr(idx), None))
        got = [int(i) for i in idx if i >= 0]
        for i in got:
            excluded[i] = 1
            curr_pfront = np.vstack([curr_pfront, Yc[i]])
        chosen.extend(got)
        if len(got) < want:  # no candidate has a positive contribution: random pick (hvi.py:38-39)
            i = int(np.random.choice(np.flatnonzero(excluded == 0)))
            excluded[i] = 1
            curr_pfront = np.vstack([curr_pfront, Yc[i]])
            chosen.append(i)
    return np.array(chosen, dtype=np.int64)


class HypervolumeImprovement(Selection):
    '''
    Selection based on Hypervolume improvement.
    '''
    def _select(self, X_candidate, Y_candidate, X, Y, batch_size):
        idx = hvi_select_indices(Y_candidate, Y, batch_size, device=getattr(self.surrogate_model, 'device', 0))
        return X_candidate[idx]
