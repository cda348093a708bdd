"""Pareto tools — same functions as the reference's `autooed/utils/pareto.py` (:12-71), with the O(n^2)
domination test and the hypervolume running on the GPU."""
from collections.abc import Iterable

import numpy as np

from autooed_b200 import _lib


def convert_minimization(Y, obj_type=None):
    if obj_type is None:
        return Y
    if isinstance(obj_type, str):
        obj_type = [obj_type] * Y.shape[1]
    assert isinstance(obj_type, Iterable), f'Objective type {type(obj_type)} is not supported'
    maxm_idx = np.array(obj_type) == 'max'
    Y = Y.copy()
    Y[:, maxm_idx] = -Y[:, maxm_idx]
    return Y


def check_pareto(Y, obj_type=None, device=0):
    """Boolean non-dominated mask (pareto.py:49-62)."""
    Y = _lib.as_f64(convert_minimization(np.asarray(Y, dtype=np.float64), obj_type))
    if Y.ndim != 2:
        raise ValueError('Y must be 2-D')
    mask = np.zeros(len(Y), dtype=np.uint8)
    if len(Y):
        _lib.require_device()
        _lib.check(_lib.load().aoed_pareto_mask(device, Y.shape[0], Y.shape[1], _lib.ptr(Y), _lib.ptr(mask)))
    return mask.astype(bool)


def find_pareto_front(Y, return_index=False, obj_type=None, device=0):
    """Non-dominated rows of Y in argsort(Y[:, 0]) order (pareto.py:27-46)."""
    if len(Y) == 0:
        return np.array([])
    Y = convert_minimization(np.asarray(Y, dtype=np.float64), obj_type)
    mask = check_pareto(Y, device=device)
    sorted_indices = np.argsort(Y.T[0])
    pareto_indices = [int(i) for i in sorted_indices if mask[i]]
    pareto_front = Y[pareto_indices].copy()
    if return_index:
        return pareto_front, pareto_indices
    return pareto_front


def pareto_rank(Y, device=0):
    """Non-dominated sorting rank (0 = first front)."""
    Y = _lib.as_f64(Y)
    rank = np.zeros(len(Y), dtype=np.int32)
    if len(Y):
        _lib.require_device()
        _lib.check(_lib.load().aoed_pareto_rank(device, Y.shape[0], Y.shape[1], _lib.ptr(Y), _lib.ptr(rank)))
    return rank


def hv_contributions(Y_candidate, front, ref_point, device=0):
    """Exclusive hypervolume of every candidate w.r.t. `front` (minimisation, bounded by ref_point)."""
    Yc = _lib.as_f64(Y_candidate)
    front = _lib.as_f64(np.asarray(front, dtype=np.float64).reshape(-1, Yc.shape[1]))
    ref = _lib.as_f64(ref_point)
    out = np.zeros(len(Yc))
    if len(Yc):
        _lib.require_device()
        _lib.check(_lib.load().aoed_hv_contributions(device, Yc.shape[0], Yc.shape[1], _lib.ptr(Yc), front.shape[0],
                                                     _lib.ptr(front), _lib.ptr(ref), _lib.ptr(out)))
    return out


def calc_hypervolume(Y, ref_point, obj_type=None, device=0):
    """Hypervolume of Y (pareto.py:65-71): accumulated exclusive contributions of the front, one point at a time."""
    Y = convert_minimization(np.asarray(Y, dtype=np.float64), obj_type)
    front = find_pareto_front(Y, device=device)
    ref = np.asarray(ref_point, dtype=np.float64)
    hv, cur = 0.0, np.empty((0, Y.shape[1]))
    for y in front:
        hv += float(hv_contributions(y[None], cur, ref, device=device)[0])
        cur = np.vstack([cur, y])
    return hv
