"""Seeded differential fuzzing: random small instances of every model family, GPU (every kernel the
dispatcher can pick, plus the generic one) against the CPU oracle, whole grid, bit for bit.  Parameters
are drawn to hit the awkward corners: non-integer costs, negative cash bounds, zero and huge penalties,
actions wider than the grid, demand tables with gaps, ties, MIN and MAX, discounting, sharding."""
import os

import numpy as np
import pytest

import sdpb200 as S
from sdpb200 import abi as A

pytestmark = pytest.mark.gpu


def _pmf(rng, T, dmax, gaps=False):
    rows = []
    for _ in range(T):
        if gaps:
            vals = np.sort(rng.choice(np.arange(0, dmax + 1), size=rng.integers(1, min(5, dmax + 1) + 1), replace=False))
        else:
            lo = rng.integers(0, 3)
            vals = np.arange(lo, lo + rng.integers(1, dmax + 1))
        p = rng.dirichlet(np.ones(len(vals)))
        if rng.random() < 0.3:
            p[rng.integers(len(p))] = 0.0          # zero-probability support point
        rows.append(np.stack([vals.astype(float), p], axis=1))
    return rows


def _cost(rng, integer):
    return float(rng.integers(0, 9)) if integer else float(rng.choice([0, 0.5, 1.25, 2, 3.7, 10]))


def make_model(family, rng):
    T = int(rng.integers(1, 4))
    if family == "A":
        return S.inventory_model(_pmf(rng, T, 7, gaps=rng.random() < 0.3), fixed_cost=_cost(rng, False),
                                 vari_cost=_cost(rng, False), hold_cost=_cost(rng, False), penalty_cost=_cost(rng, False),
                                 max_order=int(rng.integers(0, 12)), inv_min=-float(rng.integers(0, 9)),
                                 inv_max=float(rng.integers(0, 12)), direction=int(rng.integers(0, 2)),
                                 gy_mode=rng.random() < 0.2)
    if family == "B":
        lead = int(rng.integers(1, 3))
        return S.leadtime_model(_pmf(rng, T, 5), fixed_cost=_cost(rng, False), vari_cost=_cost(rng, False),
                                hold_cost=_cost(rng, False), penalty_cost=_cost(rng, False),
                                max_order=int(rng.integers(1, 7)), inv_min=-float(rng.integers(2, 9)),
                                inv_max=float(rng.integers(3, 10)), lead_time=lead, clamp=True)
    integer = family in ("Cint", "Fint")
    cash = dict(inv_min=float(rng.integers(0, 3)), inv_max=float(rng.integers(4, 12)),
                cash_min=-float(rng.integers(0, 20)), cash_max=float(rng.integers(15, 60)))
    if family in ("C", "Cint", "F", "Fint"):
        q = (A.Q_LONGDIV, 1.0, 1.0) if integer or rng.random() < 0.5 else (A.Q_DIV, 2.0, 2.0)
        kw = dict(price=_cost(rng, integer) + 1, vari_cost=_cost(rng, integer) + 1, fixed_cost=_cost(rng, integer),
                  salvage=_cost(rng, False), max_order=int(rng.integers(1, 9)), quantiser=q[0], q_mul=q[1], q_div=q[2],
                  gamma=float(rng.choice([1.0, 0.9])), direction=int(rng.integers(0, 2)), **cash)
        if not integer:
            kw.update(hold_cost=_cost(rng, False), overhead=_cost(rng, False), overhead_rate=float(rng.choice([0, 0.125])),
                      deposit_rate=float(rng.choice([0, 0.25])), penalty_cost=float(rng.choice([0, 0.5])))
        else:
            kw.update(overhead=float(rng.integers(0, 4)))
        if family in ("F", "Fint"):
            kw["recursion"] = A.REC_SURVIVAL
            kw["direction"] = A.MAX
        return S.cash_constraint_model(_pmf(rng, T, 6), **kw)
    if family == "D":
        q = [(A.Q_LONGDIV, 10.0, 10.0), (A.Q_DIV, 4.0, 4.0), (A.Q_LONGDIV, 1.0, 1.0)][int(rng.integers(0, 3))]
        return S.cash_overdraft_model(_pmf(rng, T, 6), price=_cost(rng, False) + 1, vari_cost=_cost(rng, False) + 0.5,
                                      fixed_cost=_cost(rng, False), salvage=_cost(rng, False),
                                      overhead_t=[_cost(rng, False) for _ in range(T)], r0=float(rng.choice([0, 0.125])),
                                      r2=0.125, r3=1.5, od_limit=float(rng.integers(3, 30)),
                                      interest_free=float(rng.integers(0, 3)), max_order=int(rng.integers(1, 8)),
                                      quantiser=q[0], q_mul=q[1], q_div=q[2], gamma=float(rng.choice([1.0, 0.95])), **cash)
    if family in ("DL", "DT"):
        q = [(A.Q_LONGDIV, 10.0, 10.0), (A.Q_DIV, 4.0, 4.0), (A.Q_LONGDIV, 1.0, 1.0)][int(rng.integers(0, 3))]
        kw = dict(price=_cost(rng, False) + 1, vari_cost=_cost(rng, False) + 0.5, fixed_cost=_cost(rng, False),
                  hold_cost=_cost(rng, False), interest_rate=float(rng.choice([0.0, 0.125, 0.2])),
                  max_order=int(rng.integers(1, 8)), quantiser=q[0], q_mul=q[1], q_div=q[2],
                  gamma=float(rng.choice([1.0, 0.95])), **cash)
        if family == "DL":
            return S.cash_overdraft_limit_model(_pmf(rng, T, 6), salvage=_cost(rng, False), deposit_rate=float(rng.choice([0, 0.0625])),
                                                overhead_t=[_cost(rng, False) for _ in range(T)], **kw)
        return S.cash_overdraft_testing_model(_pmf(rng, T, 6), min_cash_required=-float(rng.integers(0, 10)), **kw)
    if family == "E":
        m = S.cash_leadtime_model(_pmf(rng, T, 5), price=_cost(rng, False) + 1, vari_cost=_cost(rng, False) + 0.5,
                                  salvage=_cost(rng, False), overhead_t=[_cost(rng, False) for _ in range(T)],
                                  od_limit=float(rng.integers(3, 30)), max_order=int(rng.integers(1, 5)),
                                  q_mul=2.0, q_div=2.0, **cash)
        return m
    if family == "XR":
        return S.cash_xr_model(_pmf(rng, T, 5), price=float(rng.integers(2, 7)), vari_cost=float(rng.integers(1, 4)),
                               fixed_cost=_cost(rng, False), hold_cost=_cost(rng, False), salvage=_cost(rng, False),
                               max_order=int(rng.integers(2, 14)), inv_min=0.0, inv_max=float(rng.integers(5, 12)),
                               cash_min=-float(rng.integers(0, 8)), cash_max=float(rng.integers(15, 40)))
    if family == "M2":
        rows = []
        for _ in range(T):
            d1 = np.sort(rng.choice(np.arange(0, 5), size=rng.integers(1, 4), replace=False))
            d2 = np.sort(rng.choice(np.arange(0, 5), size=rng.integers(1, 4), replace=False))
            p1, p2 = rng.dirichlet(np.ones(len(d1))), rng.dirichlet(np.ones(len(d2)))
            rows.append(np.array([(a, b, pa * pb) for a, pa in zip(d1, p1) for b, pb in zip(d2, p2)], dtype=float))
        return S.two_product_cash_model(rows, price=(float(rng.integers(2, 8)), float(rng.integers(2, 12))),
                                        vari_cost=(float(rng.integers(1, 4)), float(rng.integers(1, 5))),
                                        salvage=(_cost(rng, False), _cost(rng, False)), q_bound=int(rng.integers(1, 5)),
                                        inv_max=float(rng.integers(3, 8)), cash_min=0.0, cash_max=float(rng.integers(10, 40)),
                                        gamma=float(rng.choice([1.0, 0.9])), tie_tolerance=float(rng.choice([0.1, 0.0])))
    raise ValueError(family)


FAMILIES = ["A", "B", "C", "Cint", "F", "Fint", "D", "DL", "DT", "E", "XR", "M2"]
KERNELS = {"A": (S.KERNEL_AUTO, S.KERNEL_GENERIC, S.KERNEL_TILED, S.KERNEL_TILED2, S.KERNEL_FUSED),
           "B": (S.KERNEL_AUTO, S.KERNEL_GENERIC, S.KERNEL_STAGED, S.KERNEL_LEAD_SLAB, S.KERNEL_LEAD_COL, S.KERNEL_LEAD_Q2),
           "C": (S.KERNEL_AUTO, S.KERNEL_GENERIC, S.KERNEL_CASH_ROW),
           "D": (S.KERNEL_AUTO, S.KERNEL_GENERIC, S.KERNEL_CASH_ROW),
           "Cint": (S.KERNEL_AUTO, S.KERNEL_GENERIC, S.KERNEL_CASH_INT),
           "Fint": (S.KERNEL_AUTO, S.KERNEL_GENERIC, S.KERNEL_CASH_INT)}


@pytest.mark.parametrize("family", FAMILIES)
def test_fuzz_family(family, oracle):
    rng = np.random.default_rng(abs(hash(family)) % (2 ** 31) if False else sum(map(ord, family)) * 7919)
    for it in range(int(os.environ.get("SDPB_FUZZ_ITERS", "40"))):  # raise for a longer one-off campaign
        spec = make_model(family, rng)
        Vo, Qo, evals, _ = oracle.dense(spec)
        for kernel in KERNELS.get(family, (S.KERNEL_AUTO, S.KERNEL_GENERIC)):
            for dedup in ((False, True) if spec.lead_time else (False,)):
                try:
                    s = S.Solver(spec, kernel=kernel, dedup=dedup)
                except S.SdpbError as e:
                    # a specifically requested kernel may not exist for this instance (documented: SDPB_ERR_ARG)
                    assert kernel not in (S.KERNEL_AUTO, S.KERNEL_GENERIC) and e.code == A.SDPB_ERR_ARG, str(e)
                    continue
                s.solve()
                for t in range(1, spec.T + 1):
                    V, Q = s.period_tables(t)
                    ok = np.array_equal(V, Vo[t - 1]) and np.array_equal(Q, Qo[t - 1])
                    assert ok, f"{family} #{it} kernel={kernel} dedup={dedup} t={t}: {spec}"
                assert s.stats()["evals"] == evals
                s.close()
