"""Model descriptors: the reference drivers' lambdas (A, f, c) lowered to `sdpb_model` parameters.

Each factory below names the reference driver whose `getFeasibleAction`, `stateTransition` and
`immediateValue` lambdas it stands for, with that driver's constants as defaults
(paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _abi as A


@dataclass
class ModelSpec:
    cost_kind: int
    pmf: Sequence[np.ndarray]            # per period: array [D_t, 2] of (demand, prob) as GetPmf returns
    inv_min: float
    inv_max: float
    max_order_idx: int
    direction: int = A.MIN
    recursion: int = A.REC_EXPECT
    lead_time: int = 0
    flags: int = A.F_CLAMP_INV
    gamma: float = 1.0
    step: float = 1.0
    cash_min: float = 0.0
    cash_max: float = 0.0
    quantiser: int = A.Q_LONGDIV
    q_mul: float = 1.0
    q_div: float = 1.0
    q_from_period: int = 0
    fixed_cost: float = 0.0
    vari_cost: float = 0.0
    hold_cost: float = 0.0
    penalty_cost: float = 0.0
    price: float = 0.0
    salvage: float = 0.0
    deposit_rate: float = 0.0
    overhead_rate: float = 0.0
    overhead: float = 0.0
    r0: float = 0.0
    r2: float = 0.0
    r3: float = 0.0
    od_limit: float = 0.0
    interest_free: float = 0.0
    price_t: Optional[Sequence[float]] = None
    vari_cost_t: Optional[Sequence[float]] = None
    overhead_t: Optional[Sequence[float]] = None
    reserve_t: Optional[Sequence[float]] = None
    reserve2: float = 0.0
    price2: float = 0.0
    vari_cost2: float = 0.0
    salvage2: float = 0.0
    tie_tolerance: float = 0.0
    apmf: Optional[Sequence] = None          # staff kind: apmf[t][y] = probabilities of turnover 0..len-1
    min_level_t: Optional[Sequence[float]] = None
    # terminal boundary function (FinalCash.BoundaryFuncton, CashRecursionV.java:125-128): V_{T+1} on the dense grid
    # in the library's state order, or None for the engines without one (Recursion.java:140)
    terminal_value: Optional[np.ndarray] = None
    # sdpb_allow bits this model is known to need (e.g. an XR instance whose order-up-to range is capped on purpose);
    # OR-ed into sdpb_options.allow by Solver / Group
    allow: int = 0
    name: str = ""

    @property
    def T(self):
        return len(self.pmf)

    @property
    def has_cash(self):
        return self.cost_kind not in (A.COST_BACKORDER, A.COST_STAFF)

    @property
    def staff(self):
        return self.cost_kind == A.COST_STAFF

    @property
    def two_product(self):
        return self.cost_kind == A.COST_CASH_TWO_PRODUCT

    @property
    def ndim(self):
        return 1 + (1 if self.two_product else 0) + (1 if self.has_cash else 0) + self.lead_time

    def to_struct(self) -> A.SdpbModel:
        """Build the C struct; the arrays it points to are kept alive on the struct object."""
        m = A.SdpbModel()
        m.struct_size = C.sizeof(A.SdpbModel)
        for f in ("cost_kind", "recursion", "direction", "lead_time", "flags", "max_order_idx", "gamma",
                  "inv_min", "inv_max", "step", "cash_min", "cash_max", "quantiser", "q_from_period", "q_mul", "q_div",
                  "fixed_cost", "vari_cost", "hold_cost", "penalty_cost", "price", "salvage",
                  "deposit_rate", "overhead_rate", "overhead", "r0", "r2", "r3", "od_limit",
                  "interest_free", "reserve2", "price2", "vari_cost2", "salvage2", "tie_tolerance"):
            setattr(m, f, getattr(self, f))
        m.T = self.T
        lens = np.ascontiguousarray([len(r) for r in self.pmf], dtype=np.int32)
        d = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float64)[:, 0] for r in self.pmf]))
        p = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float64)[:, -1] for r in self.pmf]))
        keep = [lens, d, p]
        if self.two_product:  # rows are (d1, d2, p) as GetPmfMulti.getPmf returns them
            d2 = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float64)[:, 1] for r in self.pmf]))
            keep.append(d2)
            m.pmf_d2 = d2.ctypes.data_as(C.POINTER(C.c_double))
        m.pmf_len = lens.ctypes.data_as(C.POINTER(C.c_int32))
        m.pmf_d = d.ctypes.data_as(C.POINTER(C.c_double))
        m.pmf_p = p.ctypes.data_as(C.POINTER(C.c_double))
        if self.staff:
            n_inv = int(round((self.inv_max - self.inv_min) / self.step)) + 1
            if len(self.apmf) != self.T or any(len(rows) != n_inv for rows in self.apmf):
                raise ValueError("apmf must be [T][n_inv] rows")
            alen = np.ascontiguousarray([len(r) for rows in self.apmf for r in rows], dtype=np.int32)
            ap = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float64) for rows in self.apmf for r in rows]))
            ml = np.ascontiguousarray(self.min_level_t, dtype=np.float64)
            keep += [alen, ap, ml]
            m.apmf_len = alen.ctypes.data_as(C.POINTER(C.c_int32))
            m.apmf_p = ap.ctypes.data_as(C.POINTER(C.c_double))
            m.min_level_t = ml.ctypes.data_as(C.POINTER(C.c_double))
        for f in ("price_t", "vari_cost_t", "overhead_t", "reserve_t"):
            v = getattr(self, f)
            if v is not None:
                arr = np.ascontiguousarray(v, dtype=np.float64)
                if arr.shape != (self.T,):
                    raise ValueError(f"{f} must have T={self.T} entries")
                keep.append(arr)
                setattr(m, f, arr.ctypes.data_as(C.POINTER(C.c_double)))
        if self.terminal_value is not None:
            tv = np.ascontiguousarray(self.terminal_value, dtype=np.float64).ravel()
            keep.append(tv)
            m.terminal_value = tv.ctypes.data_as(C.POINTER(C.c_double))
        m._keep = keep  # the struct owns its arrays
        return m

    # ---- the dense grid, as the library lays it out (sdpb_state_of_index) ----
    def grid_axes(self):
        """-> (inventory values, pipeline quantities, cash values): the axes of the dense grid.  Flattened order:
        inventory outermost, [second inventory,] preQ1, preQ2, cash innermost."""
        import math

        def jround(x):  # Math.round
            r = math.floor(x)
            return int(r) + (1 if x - r >= 0.5 else 0)

        n_inv = jround((self.inv_max - self.inv_min) / self.step) + 1
        inv = self.inv_min + np.arange(n_inv, dtype=np.float64) * self.step
        q = np.arange(self.max_order_idx + 1, dtype=np.float64) * self.step
        cash = None
        if self.has_cash:
            if self.cost_kind == A.COST_CASH_XR:
                raise NotImplementedError("the (x, R) grid is not tabulated on the host")

            def quantise(w):
                if self.quantiser == A.Q_TRUNC:
                    return float(int(w))
                kk = jround(w * self.q_mul)
                return kk / self.q_div if self.quantiser == A.Q_DIV else float(int(kk / int(self.q_div)) if kk >= 0
                                                                                 else -int(-kk / int(self.q_div)))

            def k_of(wq):
                return jround(wq * self.q_div) if self.quantiser == A.Q_DIV else int(wq)

            kmin, kmax = k_of(quantise(self.cash_min)), k_of(quantise(self.cash_max))
            k = np.arange(kmin, kmax + 1, dtype=np.float64)
            cash = k / self.q_div if self.quantiser == A.Q_DIV else k
        return inv, q, cash

    def tabulate(self, fn):
        """Evaluate a boundary function b(state) on every grid point, in the library's flattened order and with the
        state in API order: (inv) | (inv, preQ[, preQ2]) | (inv, cash) | (inv, cash, preQ) | (inv1, inv2, cash).
        `fn` receives numpy arrays (one per state dimension) and returns an array -- this is how a
        FinalCash.BoundaryFuncton lambda becomes `terminal_value`."""
        inv, q, cash = self.grid_axes()
        axes = [inv]
        if self.two_product:
            axes.append(inv)
        axes += [q] * self.lead_time
        if cash is not None:
            axes.append(cash)
        mesh = list(np.meshgrid(*axes, indexing="ij"))
        # API order puts the cash before the pipeline quantities
        if cash is not None and self.lead_time:
            n_lead = self.lead_time
            api = mesh[:len(mesh) - 1 - n_lead] + [mesh[-1]] + mesh[len(mesh) - 1 - n_lead:-1]
        else:
            api = mesh
        out = np.asarray(fn(*api), dtype=np.float64)
        return np.ascontiguousarray(np.broadcast_to(out, mesh[0].shape)).ravel()

    def evals_dense(self) -> float:
        """sum_t sum_s |A_t(s)| * D_t on the dense grid for state-independent action sets."""
        n_states = self.n_states()
        tot = 0.0
        for t, row in enumerate(self.pmf):
            nA = self.max_order_idx + 1
            if (self.flags & A.F_NO_ORDER_LAST) and t == self.T - 1:
                nA = 1
            tot += n_states * nA * len(row)
        return tot

    def n_states(self) -> int:
        n = int(round((self.inv_max - self.inv_min) / self.step)) + 1
        if self.two_product:
            n *= n
        n *= (self.max_order_idx + 1) ** self.lead_time
        return n  # cash axis excluded (the library reports the exact figure: sdpb_grid_info)


# ---------------------------------------------------------------------------------------------
def inventory_model(pmf, fixed_cost=100.0, vari_cost=0.0, hold_cost=1.0, penalty_cost=10.0,
                    max_order=500, inv_min=-500.0, inv_max=500.0, step=1.0, direction=A.MIN,
                    gy_mode=False, name="inventory") -> ModelSpec:
    """Single-item / capacitated stochastic lot sizing.
    Lambdas: src/capacitated/CLSPTesting.java:78-106 (same in CLSP.java:251-272,
    CLSPforDraw.java:73-103, fitss/LevelFitsS.java:74-102); G(y) pass CLSPforDraw.java:147-170.
    Engine: src/sdp/inventory/Recursion.java:89-163."""
    flags = A.F_CLAMP_INV | (A.F_GY_MODE if gy_mode else 0)
    return ModelSpec(cost_kind=A.COST_BACKORDER, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order / step), direction=direction, flags=flags, step=step,
                     fixed_cost=fixed_cost, vari_cost=vari_cost, hold_cost=hold_cost,
                     penalty_cost=penalty_cost, name=name)


def leadtime_model(pmf, fixed_cost=0.0, vari_cost=1.0, hold_cost=2.0, penalty_cost=10.0, max_order=100,
                   inv_min=-150.0, inv_max=300.0, step=1.0, lead_time=1, clamp=False,
                   name="leadtime") -> ModelSpec:
    """Lead-time inventory model.  Lambdas: src/leadtime/Leadtime.java:50-81 (lead time 1, the
    transition is NOT clamped there, :65-66).  lead_time=2 with clamp=True is the synthetic C4
    extension.  Engine: src/sdp/inventory/LeadtimeRecursion.java:47-75 (MIN only)."""
    return ModelSpec(cost_kind=A.COST_BACKORDER, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order / step), direction=A.MIN, lead_time=lead_time,
                     flags=(A.F_CLAMP_INV if clamp else 0), step=step, fixed_cost=fixed_cost,
                     vari_cost=vari_cost, hold_cost=hold_cost, penalty_cost=penalty_cost, name=name)


def cash_constraint_model(pmf, price=10.0, vari_cost=1.0, fixed_cost=0.0, hold_cost=0.0, salvage=0.5,
                          overhead=0.0, overhead_rate=0.0, deposit_rate=0.0, penalty_cost=0.0,
                          max_order=100, inv_min=0.0, inv_max=500.0, cash_min=0.0, cash_max=2000.0,
                          quantiser=A.Q_DIV, q_mul=10.0, q_div=10.0, gamma=1.0, direction=A.MAX,
                          recursion=A.REC_EXPECT, name="cash_constraint") -> ModelSpec:
    """Cash-constrained inventory.  Lambdas: src/cash/singleItem/CashConstraint.java:95-133
    (quantiser round(w*10)/10.0 at :131; CashConstraintTesting.java:146 uses round(w*1)/1).
    Engine: src/sdp/cash/CashRecursion.java:79-140 (MAX, discount factor)."""
    T = len(pmf)
    return ModelSpec(cost_kind=A.COST_CASH_DEPOSIT, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order), direction=direction, recursion=recursion,
                     flags=A.F_CLAMP_INV | A.F_LOST_SALES | A.F_CASH_LIMITED_ACTIONS, gamma=gamma,
                     cash_min=cash_min, cash_max=cash_max, quantiser=quantiser, q_mul=q_mul, q_div=q_div,
                     fixed_cost=fixed_cost, vari_cost=vari_cost, hold_cost=hold_cost,
                     penalty_cost=penalty_cost, price=price, salvage=salvage, deposit_rate=deposit_rate,
                     overhead_rate=overhead_rate, overhead=overhead,
                     reserve_t=[overhead] * T, reserve2=fixed_cost, name=name)


def cash_overdraft_model(pmf, price=10.0, vari_cost=1.0, fixed_cost=0.0, salvage=0.0, overhead_t=None,
                         r0=0.0, r2=0.1, r3=2.0, od_limit=1000.0, interest_free=0.0, max_order=100,
                         inv_min=0.0, inv_max=100.0, cash_min=-200.0, cash_max=800.0,
                         quantiser=A.Q_LONGDIV, q_mul=10.0, q_div=10.0, gamma=1.0,
                         name="cash_overdraft") -> ModelSpec:
    """Overdraft model.  Lambdas: src/cash/overdraft/CashOverdraft.java:72-118 (four-branch interest
    :88-95; quantiser round(w*10)/10 with LONG division :116).  Engine: CashRecursion (MAX)."""
    T = len(pmf)
    if overhead_t is None:
        overhead_t = [100.0] * T
    return ModelSpec(cost_kind=A.COST_CASH_OVERDRAFT, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order), direction=A.MAX,
                     flags=A.F_CLAMP_INV | A.F_LOST_SALES, gamma=gamma, cash_min=cash_min,
                     cash_max=cash_max, quantiser=quantiser, q_mul=q_mul, q_div=q_div,
                     fixed_cost=fixed_cost, vari_cost=vari_cost, price=price, salvage=salvage,
                     overhead_t=list(overhead_t), r0=r0, r2=r2, r3=r3, od_limit=od_limit,
                     interest_free=interest_free, name=name)


def cash_leadtime_model(pmf, price=5.0, vari_cost=1.0, salvage=0.5, overhead_t=None, r0=0.0, r2=0.1,
                        r3=2.0, od_limit=500.0, interest_free=0.0, max_order=30, inv_min=0.0,
                        inv_max=60.0, cash_min=-200.0, cash_max=300.0, quantiser=A.Q_DIV, q_mul=100.0,
                        q_div=100.0, name="cash_leadtime") -> ModelSpec:
    """Cash + lead time 1.  Lambdas: src/cash/overdraft/SingleProductLeadtime.java:72-119 (no order
    in the last period :74-75; quantiser round(w*100)/100.0 :117).
    Engine: src/sdp/cash/CashLeadtimeRecursion.java:48-79 (MAX, no discount)."""
    T = len(pmf)
    if overhead_t is None:
        overhead_t = [0.0] * T
    return ModelSpec(cost_kind=A.COST_CASH_OVERDRAFT, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order), direction=A.MAX, lead_time=1,
                     flags=A.F_CLAMP_INV | A.F_LOST_SALES | A.F_NO_ORDER_LAST, cash_min=cash_min,
                     cash_max=cash_max, quantiser=quantiser, q_mul=q_mul, q_div=q_div,
                     vari_cost=vari_cost, price=price, salvage=salvage, overhead_t=list(overhead_t),
                     r0=r0, r2=r2, r3=r3, od_limit=od_limit, interest_free=interest_free, name=name)


def cash_survival_model(pmf, price_t=None, vari_cost_t=None, overhead_t=None, salvage=0.5, hold_cost=0.0,
                        deposit_rate=0.0, fixed_cost=0.0, max_order=1000, inv_min=0.0, inv_max=1000.0,
                        cash_min=-500.0, cash_max=5000.0, name="cash_survival") -> ModelSpec:
    """Survival-probability model.  Lambdas: src/cash/risk/cashSurvival.java:102-147 (per-period
    price/cost arrays, quantiser round(w*1)/1 :140).  Engine: src/sdp/cash/RiskRecursion.java:64-108."""
    T = len(pmf)
    price_t = [4.0] * T if price_t is None else list(price_t)
    vari_cost_t = [1.0] * T if vari_cost_t is None else list(vari_cost_t)
    overhead_t = [100.0] * T if overhead_t is None else list(overhead_t)
    return ModelSpec(cost_kind=A.COST_CASH_DEPOSIT, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order), direction=A.MAX, recursion=A.REC_SURVIVAL,
                     flags=A.F_CLAMP_INV | A.F_LOST_SALES | A.F_CASH_LIMITED_ACTIONS,
                     cash_min=cash_min, cash_max=cash_max, quantiser=A.Q_LONGDIV, q_mul=1.0, q_div=1.0,
                     fixed_cost=fixed_cost, hold_cost=hold_cost, salvage=salvage,
                     deposit_rate=deposit_rate, price_t=price_t, vari_cost_t=vari_cost_t,
                     overhead_t=overhead_t, name=name)


def cash_xr_model(pmf, price=4.0, vari_cost=2.0, fixed_cost=0.0, hold_cost=0.0, salvage=1.0,
                  overhead=0.0, overhead_rate=0.0, deposit_rate=0.0, max_order=200, inv_min=0.0,
                  inv_max=500.0, cash_min=-100.0, cash_max=2000.0, gamma=1.0, name="cash_xr") -> ModelSpec:
    """(x, R) formulation.  Lambdas: src/cash/singleItem/CashConstraintXR.java:71-110.
    Engine: src/sdp/cash/CashRecursionXR.java:79-125.  `max_order` caps the number of order-up-to
    levels per state on the dense grid; the reference has no cap (CashConstraintXR.java:71-75), so choose it
    >= (cash_max + v*inv_max)/v - inv_min.  A cap that bites at a state the recursion visits is reported by
    sdpb_reach (SDPB_ERR_OFFGRID) unless `allow_cap`."""
    return ModelSpec(cost_kind=A.COST_CASH_XR, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order), direction=A.MAX,
                     flags=A.F_CLAMP_INV | A.F_LOST_SALES, gamma=gamma, cash_min=cash_min,
                     cash_max=cash_max, quantiser=A.Q_LONGDIV, q_mul=1.0, q_div=1.0,
                     fixed_cost=fixed_cost, vari_cost=vari_cost, hold_cost=hold_cost, price=price,
                     salvage=salvage, deposit_rate=deposit_rate, overhead_rate=overhead_rate,
                     overhead=overhead, name=name)


def cash_overdraft_limit_model(pmf, price=10.0, vari_cost=1.0, fixed_cost=0.0, hold_cost=0.0, salvage=0.0,
                               overhead_t=None, interest_rate=0.1, deposit_rate=0.0, max_order=100, inv_min=0.0,
                               inv_max=100.0, cash_min=-200.0, cash_max=800.0, quantiser=A.Q_LONGDIV, q_mul=10.0,
                               q_div=10.0, gamma=1.0, name="cash_overdraft_limit") -> ModelSpec:
    """Overdraft with one loan rate and one deposit rate on the balance after ordering, holding and
    overhead costs.  Lambdas: src/cash/overdraft/CashOverdraftLimit.java:62-99 (the cash-limited action
    bound computed at :63-64 is overwritten by maxOrderQuantity at :65, so actions are 0..maxQ)."""
    T = len(pmf)
    return ModelSpec(cost_kind=A.COST_CASH_OD_LIMIT, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order), direction=A.MAX, flags=A.F_CLAMP_INV | A.F_LOST_SALES,
                     gamma=gamma, cash_min=cash_min, cash_max=cash_max, quantiser=quantiser, q_mul=q_mul,
                     q_div=q_div, fixed_cost=fixed_cost, vari_cost=vari_cost, hold_cost=hold_cost, price=price,
                     salvage=salvage, overhead_t=list(overhead_t) if overhead_t is not None else [0.0] * T,
                     r2=interest_rate, deposit_rate=deposit_rate, name=name)


def cash_overdraft_testing_model(pmf, price=10.0, vari_cost=1.0, fixed_cost=0.0, hold_cost=0.0,
                                 interest_rate=0.1, min_cash_required=0.0, max_order=50, inv_min=0.0,
                                 inv_max=100.0, cash_min=-100.0, cash_max=400.0, quantiser=A.Q_DIV, q_mul=10.0,
                                 q_div=10.0, gamma=1.0, name="cash_overdraft_testing") -> ModelSpec:
    """Loan interest on the balance after revenue; the transition recomputes the balance itself.
    Lambdas: src/cash/overdraft/CashOverdraftTesting.java:78-120 (cash-limited actions :78-82,
    quantiser round(w*10)/10.0 :117)."""
    T = len(pmf)
    return ModelSpec(cost_kind=A.COST_CASH_OD_TESTING, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order), direction=A.MAX,
                     flags=A.F_CLAMP_INV | A.F_LOST_SALES | A.F_CASH_LIMITED_ACTIONS, gamma=gamma,
                     cash_min=cash_min, cash_max=cash_max, quantiser=quantiser, q_mul=q_mul, q_div=q_div,
                     fixed_cost=fixed_cost, vari_cost=vari_cost, hold_cost=hold_cost, price=price,
                     r2=interest_rate, reserve_t=[min_cash_required] * T, reserve2=fixed_cost, name=name)


def cash_loan_model(pmf, price=10.0, vari_cost=2.0, hold_cost=1.0, deposit_rate=0.01, loan_rate=0.15, salvage=5.0,
                    min_cash_required=-100.0, max_order=100, inv_min=0.0, inv_max=300.0, cash_min=-1000.0,
                    cash_max=1000.0, round_from_period=3, q_mul=0.0001, q_div=0.0001, gamma=1.0,
                    name="cash_loan") -> ModelSpec:
    """Deposit interest on max(w - v a, 0), loan interest on max(v a - w, 0), no holding cost in the last
    period.  Lambdas: src/cash/overdraft/TestPaper.java:75-110; the successor cash is `(int) nextCash`,
    rounded with Math.round(w*0.0001)/0.0001 first when period > 2 (:107-109)."""
    T = len(pmf)
    return ModelSpec(cost_kind=A.COST_CASH_LOAN, pmf=pmf, inv_min=inv_min, inv_max=inv_max,
                     max_order_idx=int(max_order), direction=A.MAX,
                     flags=A.F_CLAMP_INV | A.F_LOST_SALES | A.F_CASH_LIMITED_ACTIONS, gamma=gamma,
                     cash_min=cash_min, cash_max=cash_max, quantiser=A.Q_TRUNC, q_mul=q_mul, q_div=q_div,
                     q_from_period=round_from_period, vari_cost=vari_cost, hold_cost=hold_cost, price=price,
                     salvage=salvage, deposit_rate=deposit_rate, r2=loan_rate,
                     reserve_t=[min_cash_required] * T, name=name)


def two_product_cash_model(pmf, price=(4.0, 50.0), vari_cost=(2.0, 4.0), salvage=(1.0, 1.0), q_bound=100,
                           inv_max=200.0, cash_min=0.0, cash_max=10000.0, gamma=1.0, tie_tolerance=0.1,
                           name="two_product_cash") -> ModelSpec:
    """Two products sharing one cash account.  Lambdas: src/cash/multiItem/MultiItemCash.java:69-121
    (actions (i, j), 0 <= i, j < Qbound, affordable while v1 i + v2 j < cash + 0.1; `(int) nextCash`).
    Engine: src/sdp/cash/multiItem/CashRecursionMulti.java:81-116 (MAX, `> val + 0.1`).
    `pmf`: per period an array [D, 3] of (demand1, demand2, prob) as GetPmfMulti.getPmf returns."""
    return ModelSpec(cost_kind=A.COST_CASH_TWO_PRODUCT, pmf=pmf, inv_min=0.0, inv_max=inv_max,
                     max_order_idx=int(q_bound) - 1, direction=A.MAX, flags=A.F_CLAMP_INV | A.F_LOST_SALES,
                     gamma=gamma, cash_min=cash_min, cash_max=cash_max, quantiser=A.Q_TRUNC, q_mul=1.0, q_div=1.0,
                     price=price[0], vari_cost=vari_cost[0], salvage=salvage[0], price2=price[1],
                     vari_cost2=vari_cost[1], salvage2=salvage[1], tie_tolerance=tie_tolerance, name=name)


def workforce_model(turnover_rate, fix_cost=100.0, unit_vari_cost=10.0, salary=20.0, unit_penalty=80.0,
                    min_staff=None, max_hire=500, max_x=600, name="workforce") -> ModelSpec:
    """Workforce planning with binomial turnover whose pmf depends on the hire-up-to level.
    Lambdas: src/workforce/WorkforcePlanning.java:52-104 (pmf[t][y][j] = Binomial(y, turnoverRate[t]).prob(j),
    :52-68); engine: src/workforce/StaffRecursion.java:81-121 (MIN)."""
    from scipy import stats
    T = len(turnover_rate)
    if min_staff is None:
        min_staff = [40] * T
    apmf = []
    for t in range(T):
        rows = []
        for y in range(max_x + 1):
            rows.append(np.array([1.0]) if y == 0 else stats.binom(y, turnover_rate[t]).pmf(np.arange(y + 1)))
        apmf.append(rows)
    dummy = [np.array([[0.0, 1.0]]) for _ in range(T)]
    return ModelSpec(cost_kind=A.COST_STAFF, pmf=dummy, inv_min=0.0, inv_max=float(max_x), max_order_idx=int(max_hire),
                     direction=A.MIN, flags=A.F_CLAMP_INV, fixed_cost=fix_cost, vari_cost=unit_vari_cost,
                     hold_cost=salary, penalty_cost=unit_penalty, apmf=apmf, min_level_t=[float(x) for x in min_staff],
                     name=name)
