"""CPU tests of the parity oracle (oracle/sdp_oracle.cpp): the reference holds no golden vectors
for this path, so the oracle is pinned by (1) top-down == dense, bit for bit, on every visited
state, (2) an independent pure-Python transliteration of the Java loop on tiny cases,
(3) closed-form answers, (4) the frozen fixtures under tests/golden/."""
import math
import os

import numpy as np
import pytest

import cases


@pytest.mark.parametrize("case", cases.ALL, ids=lambda f: f.__name__[5:])
def test_topdown_equals_dense(case, oracle):
    spec, init = case()
    V, Q, evals, off = oracle.dense(spec, threads=2)
    rows, iv, _ = oracle.topdown(spec, init)
    nd = spec.ndim
    assert len(rows) > 0
    for r in rows:
        idx = oracle.index(spec, r[1:1 + nd])
        assert idx >= 0, f"visited state {r[:1 + nd]} is outside the dense grid"
        t = int(r[0])
        assert V[t - 1, idx] == r[-1]
        assert Q[t - 1, idx] == r[-2]
    # rows come out in the memo-map order: sorted by period first
    assert np.all(np.diff(rows[:, 0]) >= 0)


def test_dense_is_thread_count_independent(oracle):
    spec, _ = cases.case_C_rich()
    V1, Q1, e1, _ = oracle.dense(spec, threads=1)
    V4, Q4, e4, _ = oracle.dense(spec, threads=4)
    assert np.array_equal(V1, V4) and np.array_equal(Q1, Q4) and e1 == e4


# ---- (2) independent transliteration of Recursion.java:89-163 in pure Python -------------------
def _py_recursion(pmf, actions, f, c, direction_min, gamma=1.0):
    cache_v, cache_a = {}, {}
    T = len(pmf)

    def get(s):
        if s in cache_v:
            return cache_v[s]
        val = sys_float_max if direction_min else -sys_float_max
        best = 0.0
        for a in actions(s):
            q = 0.0
            for d, p in pmf[s[0] - 1]:
                q += p * c(s, a, d)
                if s[0] < T:
                    q += p * gamma * get(f(s, a, d))
            if (q < val) if direction_min else (q > val):
                val, best = q, a
        cache_v[s], cache_a[s] = val, best
        return val

    return get, cache_v, cache_a


sys_float_max = float(np.finfo(np.float64).max)


def test_python_transliteration_family_A(oracle):
    spec, init = cases.case_A_small()
    K, v, h, pi = spec.fixed_cost, spec.vari_cost, spec.hold_cost, spec.penalty_cost
    lo, hi = spec.inv_min, spec.inv_max

    def actions(s):
        return [float(i) for i in range(spec.max_order_idx + 1)]

    def f(s, a, d):  # CLSPTesting.java:89-94
        nx = s[1] + a - d
        nx = hi if nx > hi else nx
        nx = lo if nx < lo else nx
        return (s[0] + 1, nx)

    def c(s, a, d):  # CLSPTesting.java:96-106
        fixed = K if a > 0 else 0
        lvl = s[1] + a - d
        return fixed + v * a + h * max(lvl, 0) + pi * max(-lvl, 0)

    pmf = [[(float(d), float(p)) for d, p in row] for row in spec.pmf]
    get, cv, ca = _py_recursion(pmf, actions, f, c, True)
    val = get((1, 0.0))
    rows, iv, _ = oracle.topdown(spec, init)
    assert val == iv[0]
    assert len(rows) == len(cv)
    for r in rows:
        key = (int(r[0]), r[1])
        assert cv[key] == r[-1] and ca[key] == r[-2]


def test_python_transliteration_family_C(oracle):
    spec, init = cases.case_C_rich()
    T = spec.T
    P = spec

    def actions(s):  # CashConstraint.java:95-100
        maxq = int(min(P.max_order_idx, max(0, (s[2] - P.overhead - P.fixed_cost) / P.vari_cost)))
        return [float(i) for i in range(maxq + 1)]

    def c(s, a, d):  # CashConstraint.java:103-121
        revenue = P.price * min(s[1] + a, d)
        fixed = P.fixed_cost if a > 0 else 0
        variable = P.vari_cost * a
        deposite = (s[2] - fixed - variable) * (1 + P.deposit_rate)
        lvl = s[1] + a - d
        hold = P.hold_cost * max(lvl, 0)
        inc = (1 - P.overhead_rate) * revenue + deposite - hold - P.overhead - s[2]
        sal = P.salvage * max(lvl, 0) if s[0] == T else 0
        inc += sal
        end = s[2] + inc
        if end < 0:
            inc += P.penalty_cost * end
        return inc

    def jround(x):
        r = math.floor(x)
        return int(r) + (1 if x - r >= 0.5 else 0)

    def f(s, a, d):  # CashConstraint.java:123-133 with the round(w*1)/1 quantiser
        nx = max(0, s[1] + a - d)
        nw = s[2] + c(s, a, d)
        nw = P.cash_max if nw > P.cash_max else nw
        nw = P.cash_min if nw < P.cash_min else nw
        nx = P.inv_max if nx > P.inv_max else nx
        nx = P.inv_min if nx < P.inv_min else nx
        nw = float(int(jround(nw * 1) / 1)) if False else float(jround(nw * 1) // 1)
        return (s[0] + 1, float(nx), nw)

    pmf = [[(float(d), float(p)) for d, p in row] for row in spec.pmf]
    get, cv, ca = _py_recursion(pmf, actions, f, c, False, gamma=P.gamma)
    vals = [get((1, float(i[0]), float(i[1]))) for i in init]
    rows, iv, _ = oracle.topdown(spec, init)
    assert vals == list(iv)
    assert len(rows) == len(cv)
    for r in rows:
        key = (int(r[0]), r[1], r[2])
        assert cv[key] == r[-1] and ca[key] == r[-2]


def test_python_transliteration_family_B(oracle):
    """Leadtime.java:50-81 lambdas (lead time 1, transition NOT clamped) under LeadtimeRecursion.java:47-75."""
    spec, init = cases.case_B1_ref()
    K, v, h, pi = spec.fixed_cost, spec.vari_cost, spec.hold_cost, spec.penalty_cost

    def actions(s):  # Leadtime.java:50-58
        return [float(i) for i in range(spec.max_order_idx + 1)]

    def f(s, a, d):  # Leadtime.java:61-68: (t+1, x + preQ - d, a)
        return (s[0] + 1, s[1] + s[2] - d, a)

    def c(s, a, d):  # Leadtime.java:71-81
        fixed = K if a > 0 else 0
        variable = v * a
        lvl = s[1] + s[2] - d
        return fixed + variable + h * max(lvl, 0) + pi * max(-lvl, 0)

    pmf = [[(float(d), float(p)) for d, p in row] for row in spec.pmf]
    get, cv, ca = _py_recursion(pmf, actions, f, c, True)
    val = get((1, 0.0, 0.0))
    rows, iv, _ = oracle.topdown(spec, init)
    assert val == iv[0]
    assert len(rows) == len(cv)
    for r in rows:
        key = (int(r[0]), r[1], r[2])
        assert cv[key] == r[-1] and ca[key] == r[-2]


@pytest.mark.parametrize("case", [cases.case_D_small, cases.case_D_rich], ids=["D_small", "D_rich"])
def test_python_transliteration_family_D(case, oracle):
    """CashOverdraft.java:72-118 lambdas (four-branch interest, round(w*10)/10 with LONG division -- or the
    /10.0 variant) under CashRecursion.java:79-140 (MAX, discount factor)."""
    from sdpb200 import abi as A
    spec, init = case()
    P, T = spec, spec.T

    def actions(s):  # CashOverdraft.java:72-75
        return [float(i) for i in range(P.max_order_idx + 1)]

    def c(s, a, d):  # CashOverdraft.java:80-103
        revenue = P.price * min(s[1] + a, d)
        fixed = P.fixed_cost if a > 0 else 0
        variable = P.vari_cost * a
        lvl = s[1] + a - d
        before = s[2] - fixed - variable - P.overhead_t[s[0] - 1]
        if before >= 0:
            interest = -P.r0 * before
        elif before >= -P.interest_free:
            interest = 0
        elif before >= -P.od_limit:
            interest = P.r2 * (-before - P.interest_free)
        else:
            interest = P.r3 * (-before - P.od_limit) + P.r2 * (P.od_limit - P.interest_free)
        after = before - interest + revenue
        inc = after - s[2]
        inc += P.salvage * max(lvl, 0) if s[0] == T else 0
        return inc

    def jround(x):  # Math.round
        r = math.floor(x)
        return int(r) + (1 if x - r >= 0.5 else 0)

    def jdiv(a, b):  # Java long division truncates toward zero
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b > 0) else -q

    def f(s, a, d):  # CashOverdraft.java:106-118
        nx = max(0, s[1] + a - d)
        nw = s[2] + c(s, a, d)
        nw = P.cash_max if nw > P.cash_max else nw
        nw = P.cash_min if nw < P.cash_min else nw
        nx = P.inv_max if nx > P.inv_max else nx
        nx = P.inv_min if nx < P.inv_min else nx
        k = jround(nw * P.q_mul)
        nw = float(jdiv(k, int(P.q_div))) if P.quantiser == A.Q_LONGDIV else k / P.q_div
        return (s[0] + 1, float(nx), nw)

    pmf = [[(float(d), float(p)) for d, p in row] for row in spec.pmf]
    get, cv, ca = _py_recursion(pmf, actions, f, c, False, gamma=P.gamma)
    vals = [get((1, float(i[0]), float(i[1]))) for i in init]
    rows, iv, _ = oracle.topdown(spec, init)
    assert vals == list(iv)
    assert len(rows) == len(cv)
    for r in rows:
        key = (int(r[0]), r[1], r[2])
        assert cv[key] == r[-1] and ca[key] == r[-2]


def test_python_transliteration_family_F(oracle):
    """cashSurvival.java:102-147 lambdas under RiskRecursion.getSurvProb (RiskRecursion.java:64-108): survival
    indicator in the last period, successors with negative cash count as ruin without being expanded."""
    spec, init = cases.case_F_small()
    P, T = spec, spec.T
    pmf = [[(float(d), float(p)) for d, p in row] for row in spec.pmf]

    def actions(s):  # cashSurvival.java:102-109 (bankruptBefore states are never expanded)
        maxq = max(min(s[2] / P.vari_cost_t[s[0] - 1], P.max_order_idx), 0)
        return [float(i) for i in range(int(maxq) + 1)]

    def c(s, a, d):  # cashSurvival.java:115-128
        t = s[0] - 1
        revenue = P.price_t[t] * min(s[1] + a, d)
        fixed = P.fixed_cost if a > 0 else 0
        variable = P.vari_cost_t[t] * a
        deposite = (s[2] - fixed - variable) * (1 + P.deposit_rate)
        lvl = s[1] + a - d
        hold = P.hold_cost * max(lvl, 0)
        inc = revenue + deposite - hold - P.overhead_t[t] - s[2]
        inc += P.salvage * max(lvl, 0) if s[0] == T else 0
        return inc

    def jround(x):
        r = math.floor(x)
        return int(r) + (1 if x - r >= 0.5 else 0)

    def f(s, a, d):  # cashSurvival.java:131-146
        nx = max(0, s[1] + a - d)
        nw = s[2] + c(s, a, d)
        nw = P.cash_max if nw > P.cash_max else nw
        nw = P.cash_min if nw < P.cash_min else nw
        nx = P.inv_max if nx > P.inv_max else nx
        nx = P.inv_min if nx < P.inv_min else nx
        return (s[0] + 1, float(nx), float(jround(nw * 1) // 1))

    cv, ca = {}, {}

    def surv(s):  # RiskRecursion.java:64-108
        if s in cv:
            return cv[s]
        val, best = -sys_float_max, 0.0
        for a in actions(s):
            q = 0.0
            for d, p in pmf[s[0] - 1]:
                if s[0] == T:
                    q += p * (1 if s[2] + c(s, a, d) >= 0 else 0)
                else:
                    n = f(s, a, d)
                    q += p * (0 if n[2] < 0 else surv(n))
            if q > val:
                val, best = q, a
        cv[s], ca[s] = val, best
        return val

    vals = [surv((1, float(i[0]), float(i[1]))) for i in init]
    rows, iv, _ = oracle.topdown(spec, init)
    assert vals == list(iv)
    assert len(rows) == len(cv)
    for r in rows:
        key = (int(r[0]), r[1], r[2])
        assert cv[key] == r[-1] and ca[key] == r[-2]


def test_python_transliteration_family_XR(oracle):
    """CashConstraintXR.java:68-110 lambdas under CashRecursionXR.java:79-125: state (t, x, R = cash + v x), the
    action is the order-up-to level.  The dense grid caps the number of levels per state (`max_order_idx`,
    DESIGN.md a11); the transliteration applies the same cap and counts how often it binds."""
    spec, init = cases.case_XR_small()
    spec.max_order_idx = 46  # >= cash_max / v + 1 levels: the cap can never bind, as in the uncapped reference
    P, T = spec, spec.T
    binds = [0]

    def actions(s):  # CashConstraintXR.java:69-73
        max_y = s[1] if s[2] / P.vari_cost < s[1] else s[2] / P.vari_cost
        length = int(max_y - s[1]) + 1
        if length > P.max_order_idx + 1:
            binds[0] += 1
            length = P.max_order_idx + 1
        return [s[1] + float(i) for i in range(length)]

    def c(s, y, d):  # CashConstraintXR.java:76-89
        revenue = P.price * min(y, d)
        action = y - s[1]
        fixed = P.fixed_cost if y > s[1] else 0
        variable = P.vari_cost * action
        init_cash = s[2] - P.vari_cost * s[1]
        deposite = (init_cash - fixed - variable) * (1 + P.deposit_rate)
        lvl = y - d
        hold = P.hold_cost * max(lvl, 0)
        inc = (1 - P.overhead_rate) * revenue + deposite - hold - P.overhead - init_cash
        inc += P.salvage * max(lvl, 0) if s[0] == T else 0
        return inc

    def jround(x):
        r = math.floor(x)
        return int(r) + (1 if x - r >= 0.5 else 0)

    def f(s, y, d):  # CashConstraintXR.java:92-109
        nx = max(0, y - d)
        init_cash = s[2] - P.vari_cost * s[1]
        nw = init_cash + c(s, y, d)
        nw = P.cash_max if nw > P.cash_max else nw
        nw = P.cash_min if nw < P.cash_min else nw
        nx = P.inv_max if nx > P.inv_max else nx
        nx = P.inv_min if nx < P.inv_min else nx
        nw = float(jround(nw * 1) // 1)
        return (s[0] + 1, float(nx), nw + P.vari_cost * nx)

    pmf = [[(float(d), float(p)) for d, p in row] for row in spec.pmf]
    get, cv, ca = _py_recursion(pmf, actions, f, c, False, gamma=P.gamma)
    vals = [get((1, float(i[0]), float(i[1]))) for i in init]
    rows, iv, _ = oracle.topdown(spec, init)
    assert vals == list(iv)
    assert len(rows) == len(cv)
    for r in rows:
        key = (int(r[0]), r[1], r[2])
        assert cv[key] == r[-1] and ca[key] == r[-2]
    assert binds[0] == 0  # the cap never bound on this instance: the comparison is with the uncapped reference


# ---- (3) closed forms ---------------------------------------------------------------------------
def test_closed_form_twopoint(oracle):
    # demand 4 or 10 (p = 1/2), h = pi = 1, no ordering cost: any order-up-to level in [4,10] costs
    # 3 per period, so V_1(0) = 3*T = 9 and the first optimum in ascending order is a = 4.
    spec, init = cases.case_A_twopoint()
    rows, iv, _ = oracle.topdown(spec, init)
    assert iv[0] == 9.0
    first = rows[(rows[:, 0] == 1) & (rows[:, 1] == 0.0)][0]
    assert first[-2] == 4.0


def test_closed_form_deterministic_demand(oracle, S):
    # one demand point d = 5 with probability 1, T = 1: V(x) = min_a [K 1(a>0) + h (x+a-5)^+ + pi (5-x-a)^+]
    pmf = [np.array([[5.0, 1.0]])]
    spec = S.inventory_model(pmf, fixed_cost=3, vari_cost=0, hold_cost=1, penalty_cost=2, max_order=10,
                             inv_min=-5, inv_max=10)
    V, Q, evals, _ = oracle.dense(spec)
    for i, x in enumerate(np.arange(-5, 11)):
        cands = [(3 if a > 0 else 0) + max(x + a - 5, 0) + 2 * max(5 - x - a, 0) for a in range(11)]
        assert V[0, i] == min(cands)
        assert Q[0, i] == int(np.argmin(cands))
    assert evals == 16 * 11 * 1


def test_java_round_and_long_division(oracle, S):
    # Math.round is round-half-up; round(w*10)/10 with long division truncates toward zero
    # (CashOverdraft.java:116): cash -0.4 -> round(-4.0) = -4 -> -4/10 = 0
    pmf = [np.array([[0.0, 1.0]])] * 2
    spec = S.cash_overdraft_model(pmf, price=1, vari_cost=0.25, overhead_t=[0.4, 0.0], r2=0, r3=0,
                                  max_order=2, inv_min=0, inv_max=4, cash_min=-10, cash_max=10)
    c, nxt = oracle.eval_triple(spec, 1, [0.0, 0.0], 0.0, 0.0)
    assert c == -0.4 and nxt[1] == 0.0
    # a = 2 costs 0.5 more: cash -0.9 -> round(-9.0) = -9 -> -9/10 = 0 (truncation, not floor)
    c, nxt = oracle.eval_triple(spec, 1, [0.0, 0.0], 2.0, 0.0)
    assert c == -0.9 and nxt[1] == 0.0
    # half-up: 2.5 -> 3, -2.5 -> -2
    spec2 = S.cash_constraint_model(pmf, price=0, vari_cost=1, salvage=0, max_order=0, inv_min=0, inv_max=1,
                                    cash_min=-10, cash_max=10, quantiser=S.Q_LONGDIV, q_mul=1.0, q_div=1.0,
                                    overhead=0.5)
    _, nxt = oracle.eval_triple(spec2, 1, [0.0, 3.0], 0.0, 0.0)
    assert nxt[1] == 3.0  # 3 - 0.5 = 2.5 -> 3
    _, nxt = oracle.eval_triple(spec2, 1, [0.0, -2.0], 0.0, 0.0)
    assert nxt[1] == -2.0  # -2.5 -> -2


def test_cash_limited_action_count(oracle):
    spec, _ = cases.case_C_rich()  # v = 2, overhead 2, K = 3, maxQ 9
    assert oracle.n_actions(spec, 1, [0.0, 4.0]) == 1       # (4-2-3)/2 < 0 -> {0}
    assert oracle.n_actions(spec, 1, [0.0, 12.0]) == 4      # (12-5)/2 = 3.5 -> (int) 3 -> 4 actions
    assert oracle.n_actions(spec, 1, [0.0, 90.0]) == 10     # capped at maxQ


# ---- (4) frozen fixtures ------------------------------------------------------------------------
@pytest.mark.parametrize("name", cases.GOLDEN)
def test_oracle_reproduces_golden(name, oracle):
    spec, init, g = cases.load_golden(name)
    V, Q, evals, _ = oracle.dense(spec, threads=2)
    per = g["periods"] - 1
    assert np.array_equal(V[per], g["V"])
    assert np.array_equal(Q[per].astype(np.float32), g["Q"])
    assert evals == float(g["evals"])
    rows, iv, _ = oracle.topdown(spec, init)
    assert np.array_equal(rows, g["topdown_rows"])
    assert np.array_equal(iv, g["init_values"])


# ---- (5) the reference author's own recorded output ----------------------------------------------
def test_oracle_loop_reproduces_reference_output(oracle):
    """The SAME solve_state / TopDown code every parity test relies on (oracle/sdp_oracle.cpp), driven
    with the reference's two-product lead-time lambdas, reproduces the program output recorded in
    src/cash/overdraft/MultiProductLeadtime.java:34-44 to the last digit:
        T = 2, demands {20,30,40} x {10,15,20}, p = {.25,.5,.25}:  -17.800000000000008, Q1 = 40, Q2 = 20.
    (The T = 3 record of the same file, demands {10,30} x {5,15}: -76.56, Q1 = 30, Q2 = 15, also reproduces
    — 1.7e11 evaluations, 34 CPU-minutes; run it with SDPB_SLOW_TESTS=1.)"""
    v, q1, q2, ns = oracle.multi_lead(2, 45, [20, 30, 40], [.25, .5, .25], [10, 15, 20], [.25, .5, .25])
    assert repr(v) == "-17.800000000000008" and (q1, q2) == (40, 20)
    v, q1, q2, ns = oracle.multi_lead(2, 50, [20, 30, 40], [.25, .5, .25], [10, 15, 20], [.25, .5, .25])
    assert repr(v) == "-17.800000000000008" and (q1, q2) == (40, 20)
    if os.environ.get("SDPB_SLOW_TESTS") == "1":
        v, q1, q2, ns = oracle.multi_lead(3, 50, [10, 30], [.5, .5], [5, 15], [.5, .5])
        assert repr(v) == "-76.56" and (q1, q2) == (30, 15)


def test_reference_known_answer_two_product_T2(oracle):
    """src/cash/overdraft/MultiProductLeadtime.java:34-44 records, for discrete demands
    {20,30,40} x {10,15,20} with probabilities {.25,.5,.25} and T = 2:
        "final optimal cash  is -17.800000000000008 ... Q1 = 40, Q2 = 20".
    oracle/multi_item_oracle.cpp is the same kind of line-by-line restatement as sdp_oracle.cpp (same
    loop, same p*gamma*V association, same overdraft-interest branches, same clamp / (int) idioms) and
    reproduces that output to the last digit, for both values Qbound has had in the file (45, 50)."""
    import subprocess
    subprocess.run(["make", "-C", oracle.ORACLE_DIR, "multi_item_oracle"], check=True, capture_output=True)
    exe = os.path.join(oracle.ORACLE_DIR, "multi_item_oracle")
    for qbound in (45, 50):
        out = subprocess.run([exe, "2", str(qbound), "3", "20", "30", "40", "0.25", "0.5", "0.25",
                              "10", "15", "20", "0.25", "0.5", "0.25"], capture_output=True, text=True, check=True)
        val, q1, q2 = out.stdout.split()[:3]
        assert val == "-17.800000000000008"
        assert (int(q1), int(q2)) == (40, 20)
