"""Descriptor spot checks: evaluate (c, f, |A|) for random triples ON THE DEVICE (sdpb_eval_triples,
the same device code the solve runs) so a host facade can compare them with user lambdas."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A


def random_triple(spec, rng):
    """((period, *state), action, demand) with the state on the grid and the demand from the pmf."""
    t = rng.randint(1, spec.T)
    n_inv = int(round((spec.inv_max - spec.inv_min) / spec.step)) + 1
    x = spec.inv_min + rng.randrange(n_inv) * spec.step
    vec = [x]
    if spec.has_cash:
        if spec.cost_kind == A.COST_CASH_XR:
            vec.append(float(rng.randint(int(spec.cash_min), int(spec.cash_max))))
        else:
            lo, hi = int(np.ceil(spec.cash_min * spec.q_mul)), int(np.floor(spec.cash_max * spec.q_mul))
            if spec.quantiser == A.Q_TRUNC:
                vec.append(float(rng.randint(int(spec.cash_min), int(spec.cash_max))))
            else:
                kk = rng.randint(lo, hi)
                vec.append(kk / spec.q_div if spec.quantiser == A.Q_DIV else float(int(kk / int(spec.q_div))))
    for _ in range(spec.lead_time):
        vec.append(rng.randrange(spec.max_order_idx + 1) * spec.step)
    row = np.asarray(spec.pmf[t - 1])
    d = float(row[rng.randrange(len(row)), 0])
    ai = rng.randrange(spec.max_order_idx + 1)
    a = ai * spec.step + (x if spec.cost_kind == A.COST_CASH_XR else 0.0)
    return (t, *vec), a, d


def eval_descriptor(spec, st, a, d, solver):
    """-> (c, next state vector (API order), |A(s)|), computed by the GPU library on `solver`'s handle (the engine's
    own: no second copy of the grid, nothing cached across engines)."""
    s = solver
    t, vec = st[0], np.ascontiguousarray([st[1:]], dtype=np.float64)
    base = vec[0, 0] if spec.cost_kind == A.COST_CASH_XR else 0.0
    ai = np.ascontiguousarray([int(round((a - base) / spec.step))], dtype=np.int32)
    dem = np.ascontiguousarray([d], dtype=np.float64)
    c = np.empty(1)
    nxt = np.empty((1, s.ndim))
    na = np.empty(1, dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    s._check(s.lib.sdpb_eval_triples(s.h, t, vec.ctypes.data_as(dp), ai.ctypes.data_as(ip),
                                     dem.ctypes.data_as(dp), 1, c.ctypes.data_as(dp),
                                     nxt.ctypes.data_as(dp), na.ctypes.data_as(ip)))
    return float(c[0]), tuple(nxt[0]), int(na[0])
