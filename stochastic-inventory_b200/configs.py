"""The five measurement configurations of BASELINE.json, pinned as in SURVEY.md §8(d).

Every input is deterministic (no RNG).  `scale` arguments shrink a config for parity tests; the
defaults are the full sizes the bench uses."""
from __future__ import annotations

from . import _abi as A
from .getpmf import poisson_pmf
from .models import cash_constraint_model, inventory_model, leadtime_model

# row 4 of the demand table in src/capacitated/fitss/OneLevelFitsSTest.java:33-43
C2_MEANS = [36.3, 30, 23.7, 21, 23.7, 30, 36.3, 39, 36.3, 30, 23.7, 21, 23.7, 30, 36.3, 30.9, 24.3, 21.3, 26.4, 33]


def c1():
    """src/sdp single-item stochastic lot sizing: T=4, Poisson [20,40,60,40], K=100, h=1, pi=10,
    inventory -500..500, actions 0..500 (CLSPTesting.java:52).  Initial state x0 = 0."""
    return inventory_model(poisson_pmf([20, 40, 60, 40]), fixed_cost=100, vari_cost=0, hold_cost=1,
                           penalty_cost=10, max_order=500, inv_min=-500, inv_max=500, name="C1")


def c2():
    """src/capacitated: T=20, capacity 100, 2001 inventory states x 101 actions, K=500, h=1, pi=10
    (fitss/LevelFitsS.java:38-39 grid)."""
    return inventory_model(poisson_pmf(C2_MEANS), fixed_cost=500, vari_cost=0, hold_cost=1,
                           penalty_cost=10, max_order=100, inv_min=-1000, inv_max=1000, name="C2")


def c3(T=12, inv_max=500, cash_max=2000, max_order=200, mean=151.0):
    """src/cash: 2-D (inventory, cash) ~1e6 states, T=12, 200-point pmf.  CashConstraint.java:95-133
    lambdas, integer cash grid (quantiser round(w*1)/1 as CashConstraintTesting.java:146).
    Initial state (x0, w0) = (0, 100)."""
    return cash_constraint_model(poisson_pmf([mean] * T), price=10, vari_cost=1, fixed_cost=0, hold_cost=0,
                                 salvage=0.5, overhead=0, max_order=max_order, inv_min=0, inv_max=inv_max,
                                 cash_min=0, cash_max=cash_max, quantiser=A.Q_LONGDIV, q_mul=1.0, q_div=1.0,
                                 name="C3")


def c4(T=20, inv_half=500, max_order=100, mean=10.0):
    """src/leadtime extended to lead time 2: state (x, q1, q2), ~1e7 states, Poisson(10) -> 25 points,
    K=0, v=1, h=2, pi=10 (Leadtime.java:33-39), clamp added.  Initial state (0, 0, 0)."""
    return leadtime_model(poisson_pmf([mean] * T), fixed_cost=0, vari_cost=1, hold_cost=2, penalty_cost=10,
                          max_order=max_order, inv_min=-inv_half, inv_max=inv_half, lead_time=2, clamp=True,
                          name="C4")


def c5(n_states=10_000_000, T=4, n_actions=200, mean=151.0):
    """Synthetic scale sweep: S states x 200 actions x 200 demand points, family A, K=100, h=1, pi=10."""
    half = n_states // 2
    return inventory_model(poisson_pmf([mean] * T), fixed_cost=100, vari_cost=0, hold_cost=1, penalty_cost=10,
                           max_order=n_actions - 1, inv_min=-half, inv_max=n_states - half - 1,
                           name=f"C5_S{n_states}")
