"""ctypes access to the CPU oracle (oracle/libsdp_oracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsdp_oracle.so")

_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(ORACLE_DIR, "sdp_oracle.cpp")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        build()
    lib = C.CDLL(ORACLE_SO)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.oracle_grid.argtypes = [vp, C.POINTER(C.c_int64), ip]
    lib.oracle_dense.argtypes = [vp, dp, dp, dp, C.POINTER(C.c_int64), C.c_int]
    lib.oracle_topdown.argtypes = [vp, dp, C.c_int, dp, C.c_int64, dp, dp]
    lib.oracle_topdown.restype = C.c_int64
    lib.oracle_step_states.argtypes = [vp, C.c_int, dp, C.POINTER(C.c_int64), C.c_int, dp, dp]
    lib.oracle_multi_lead.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, dp, ip, ip,
                                      C.POINTER(C.c_int64)]
    lib.oracle_reached.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_double, C.c_double, dp, dp, dp, dp, C.POINTER(C.c_int64), ip,
                                   C.c_double, C.c_double, C.c_double, C.c_double]
    lib.oracle_simulate.argtypes = [vp, dp, dp, dp, C.c_int, C.c_double, dp]
    lib.oracle_eval.argtypes = [vp, C.c_int, dp, C.c_double, C.c_double, dp, dp]
    lib.oracle_index.argtypes = [vp, dp]
    lib.oracle_index.restype = C.c_int64
    lib.oracle_n_actions.argtypes = [vp, C.c_int, dp]
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def grid(spec):
    lib = load()
    m = spec.to_struct()
    n, nd = C.c_int64(), C.c_int()
    lib.oracle_grid(C.byref(m), C.byref(n), C.byref(nd))
    return n.value, nd.value


def dense(spec, threads=0):
    """-> V[T,S], Q[T,S], evals, offgrid_count"""
    lib = load()
    m = spec.to_struct()
    S, _ = grid(spec)
    V = np.empty((spec.T, S))
    Q = np.empty((spec.T, S))
    ev, off = C.c_double(), C.c_int64()
    lib.oracle_dense(C.byref(m), _dp(V), _dp(Q), C.byref(ev), C.byref(off), threads)
    return V, Q, ev.value, off.value


def topdown(spec, init_states):
    """-> rows [n, ndim+3] = [t, state..., Q, V] sorted by the memo order; init values; evals"""
    lib = load()
    m = spec.to_struct()
    st = np.ascontiguousarray(np.atleast_2d(np.asarray(init_states, dtype=np.float64)))
    _, nd = grid(spec)
    iv = np.empty(st.shape[0])
    ev = C.c_double()
    n = lib.oracle_topdown(C.byref(m), _dp(st), st.shape[0], None, 0, _dp(iv), C.byref(ev))
    rows = np.empty((n, nd + 3))
    lib.oracle_topdown(C.byref(m), _dp(st), st.shape[0], _dp(rows), n, _dp(iv), C.byref(ev))
    return rows, iv, ev.value


def step_states(spec, period, Vnext, idx):
    """Recompute (V_t, Q_t) at flattened indices `idx` from the full V_{t+1} table."""
    lib = load()
    m = spec.to_struct()
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    v = np.empty(len(idx))
    q = np.empty(len(idx))
    vn = None if Vnext is None else _dp(np.ascontiguousarray(Vnext, dtype=np.float64))
    lib.oracle_step_states(C.byref(m), period, vn, idx.ctypes.data_as(C.POINTER(C.c_int64)), len(idx),
                           _dp(v), _dp(q))
    return v, q


def multi_lead(T, qbound, vals1, probs1, vals2, probs2, overhead=100.0):
    """The reference's two-product lead-time model through the oracle's own loop -> (value, Q1, Q2, n_states)."""
    lib = load()
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (vals1, probs1, vals2, probs2)]
    v, q1, q2, ns = C.c_double(), C.c_int(), C.c_int(), C.c_int64()
    lib.oracle_multi_lead(T, qbound, len(vals1), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), overhead,
                          C.byref(v), C.byref(q1), C.byref(q2), C.byref(ns))
    return v.value, q1.value, q2.value, ns.value


def simulate(spec, Q, init_state, samples, discount=1.0):
    lib = load()
    m = spec.to_struct()
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    st = np.ascontiguousarray(init_state, dtype=np.float64)
    sm = np.ascontiguousarray(samples, dtype=np.float64)
    vals = np.empty(sm.shape[0])
    lib.oracle_simulate(C.byref(m), _dp(Q), _dp(st), _dp(sm), sm.shape[0], float(discount), _dp(vals))
    return vals


def eval_triple(spec, period, state, action, demand):
    lib = load()
    m = spec.to_struct()
    st = np.ascontiguousarray(state, dtype=np.float64)
    c = C.c_double()
    nxt = np.empty(len(st))
    lib.oracle_eval(C.byref(m), period, _dp(st), float(action), float(demand), C.byref(c), _dp(nxt))
    return c.value, nxt


def index(spec, state):
    lib = load()
    m = spec.to_struct()
    st = np.ascontiguousarray(state, dtype=np.float64)
    return lib.oracle_index(C.byref(m), _dp(st))


def n_actions(spec, period, state):
    lib = load()
    m = spec.to_struct()
    st = np.ascontiguousarray(state, dtype=np.float64)
    return lib.oracle_n_actions(C.byref(m), period, _dp(st))


def reached(kind, pmf, q_bound, price, vari_cost, salvage, init, deposit_rate=0.0, min_inv=0.0, max_inv=200.0, min_cash=0.0,
            max_cash=10000.0, gamma=1.0, fixed_cost=0.0, hold_cost=0.0, min_cash_required=0.0, state_q=0.1):
    """CashRecursionMultiXR (kind 1) / CashRecursionV (kind 2) / CashRecursion with the CashConstraintTest lambdas (kind 3:
    one product, pmf rows (demand, prob), init = (x, 0, cash)) through the oracle's literal top-down restatement
    -> (value, action1, action2, visited states)."""
    lib = load()
    lens = np.ascontiguousarray([len(r) for r in pmf], dtype=np.int32)
    tab = np.zeros((len(pmf), int(lens.max()), 3))
    for t, r in enumerate(pmf):
        r = np.asarray(r, dtype=np.float64)
        if r.shape[1] == 2:
            r = np.column_stack([r[:, 0], np.zeros(len(r)), r[:, 1]])
        tab[t, :len(r)] = r
    d1, d2, p = (np.ascontiguousarray(tab[:, :, k]).ravel() for k in range(3))
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (price, vari_cost, salvage, init)]
    v, a1, a2, ns = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
    lib.oracle_reached(kind, tab.shape[0], q_bound, tab.shape[1], _dp(d1), _dp(d2), _dp(p), _dp(arrs[0]), _dp(arrs[1]),
                       _dp(arrs[2]), deposit_rate, min_inv, max_inv, min_cash, max_cash, gamma, _dp(arrs[3]),
                       C.byref(v), C.byref(a1), C.byref(a2), C.byref(ns), lens.ctypes.data_as(C.POINTER(C.c_int)),
                       fixed_cost, hold_cost, min_cash_required, state_q)
    return v.value, a1.value, a2.value, ns.value
