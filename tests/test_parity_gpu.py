"""GPU parity tests: libsdpb200.so (through the C-ABI) against the CPU oracle, bit for bit.

Values are compared with ==, order quantities with == (the argopt reduction reproduces the
reference's first-wins tie rule, so even exact ties must agree)."""
import ctypes as C

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def _solve_all(S, spec, **kw):
    s = S.Solver(spec, **kw)
    s.solve()
    V = np.empty((spec.T, s.n_states))
    Q = np.empty((spec.T, s.n_states))
    for t in range(1, spec.T + 1):
        V[t - 1], Q[t - 1] = s.period_tables(t)
    return s, V, Q


@pytest.mark.parametrize("case", cases.ALL, ids=lambda f: f.__name__[5:])
def test_generic_kernel_whole_grid(case, S, oracle):
    spec, init = case()
    Vo, Qo, evals, _ = oracle.dense(spec)
    s, V, Q = _solve_all(S, spec, kernel=S.KERNEL_GENERIC)
    assert s.n_states == Vo.shape[1]
    assert np.array_equal(V, Vo)
    assert np.array_equal(Q, Qo)
    st = s.stats()
    assert st["evals"] == evals
    assert st["launches"] == spec.T and st["kernel_used"] == S.KERNEL_GENERIC


@pytest.mark.parametrize("case", cases.ALL, ids=lambda f: f.__name__[5:])
def test_auto_kernel_whole_grid(case, S, oracle):
    """Whatever kernel AUTO picks (tiled where one exists) must give the same bits."""
    spec, init = case()
    Vo, Qo, evals, _ = oracle.dense(spec)
    s, V, Q = _solve_all(S, spec)
    assert np.array_equal(V, Vo)
    assert np.array_equal(Q, Qo)
    assert s.stats()["evals"] == evals


A_CASES = [cases.case_A_small, cases.case_A_max, cases.case_A_gy, cases.case_A_twopoint, cases.case_A_halfstep,
           cases.case_A_sparse_pmf, cases.case_A_degenerate, cases.case_A_one_state]


@pytest.mark.parametrize("case", A_CASES, ids=lambda f: f.__name__[5:])
@pytest.mark.parametrize("kernel", ["tiled", "tiled2"])
def test_both_tiled_variants_whole_grid(case, kernel, S, oracle):
    """bi_inv_tiled (1 level x 8 actions per thread) and bi_inv_tiled2 (4 levels x 4 actions, sliding
    register window), forced on small grids so the action-split / last-CTA merge path runs too."""
    spec, init = case()
    Vo, Qo, evals, _ = oracle.dense(spec)
    k = S.KERNEL_TILED if kernel == "tiled" else S.KERNEL_TILED2
    s, V, Q = _solve_all(S, spec, kernel=k)
    assert np.array_equal(V, Vo)
    assert np.array_equal(Q, Qo)
    used = s.stats()["kernel_used"]
    if kernel == "tiled2" and case is not cases.case_A_twopoint:   # two-point demand is not consecutive
        assert used == S.KERNEL_TILED2
    else:
        assert used == S.KERNEL_TILED


def test_tiled2_wide_cases(S, oracle):
    """Shapes that exercise tile edges of the 1021-state tile: several tiles, D not a multiple of 4,
    action count not a multiple of 4, MAX direction."""
    for (n_inv, max_order, means, direction) in [(2500, 37, [7, 9, 8], S.MIN), (1021 * 2 + 5, 10, [3, 11], S.MAX),
                                                 (1500, 3, [2, 2, 2], S.MIN)]:
        spec = S.inventory_model(S.poisson_pmf(means, 0.999), fixed_cost=12, vari_cost=1, hold_cost=1,
                                 penalty_cost=6, max_order=max_order, inv_min=-(n_inv // 2),
                                 inv_max=n_inv - n_inv // 2 - 1, direction=direction)
        Vo, Qo, _, _ = oracle.dense(spec)
        for k in (S.KERNEL_TILED2, S.KERNEL_TILED):
            s, V, Q = _solve_all(S, spec, kernel=k)
            assert np.array_equal(V, Vo) and np.array_equal(Q, Qo), (n_inv, max_order, k)


LEAD_CASES = [cases.case_B1_ref, cases.case_B1_fixed, cases.case_B2_small, cases.case_E_small]


@pytest.mark.parametrize("case", LEAD_CASES, ids=lambda f: f.__name__[5:])
@pytest.mark.parametrize("kernel", ["auto", "generic"])
def test_dedup_whole_grid(case, kernel, S, oracle):
    """Lead-time models depend on (x, preQ) only through x + preQ; folding those states is exact."""
    spec, init = case()
    Vo, Qo, evals, _ = oracle.dense(spec)
    k = S.KERNEL_AUTO if kernel == "auto" else S.KERNEL_GENERIC
    s, V, Q = _solve_all(S, spec, dedup=True, kernel=k)
    assert np.array_equal(V, Vo)
    assert np.array_equal(Q, Qo)
    st = s.stats()
    assert st["evals"] == evals and 0 < st["evals_executed"] < evals


@pytest.mark.parametrize("case", [cases.case_C_int, cases.case_C_int_K, cases.case_F_small],
                         ids=lambda f: f.__name__[5:])
def test_integer_cash_kernels_whole_grid(case, S, oracle):
    """bi_cash_diag (AUTO) and bi_cash_int (requested) against the oracle."""
    spec, _ = case()
    Vo, Qo, _, _ = oracle.dense(spec)
    for k, used in ((S.KERNEL_AUTO, S.KERNEL_CASH_DIAG), (S.KERNEL_CASH_INT, S.KERNEL_CASH_INT)):
        s, V, Q = _solve_all(S, spec, kernel=k)
        assert s.stats()["kernel_used"] == used, spec.name
        assert np.array_equal(V, Vo) and np.array_equal(Q, Qo), (spec.name, k)


def test_cash_diag_wide_cases(S, oracle):
    """Diagonal-window kernel edges: price larger than the cash tile step, inventory rows not a multiple
    of 8, short and long demand tables (stock-out region dominant / absent), MIN direction, discounting."""
    for (price, v, K, means, inv_max, cash_max, max_order, direction, gamma) in [
            (10, 1, 0, [5, 6, 5], 21, 300, 13, S.MAX, 1.0), (3, 2, 4, [9, 3], 30, 150, 20, S.MAX, 0.9),
            (25, 3, 0, [2, 2, 2], 9, 400, 6, S.MIN, 1.0), (7, 1, 2, [30, 28], 70, 260, 40, S.MAX, 1.0)]:
        spec = S.cash_constraint_model(S.poisson_pmf(means, 0.999), price=price, vari_cost=v, fixed_cost=K,
                                       salvage=0.5, max_order=max_order, inv_min=0, inv_max=inv_max, cash_min=-20,
                                       cash_max=cash_max, quantiser=S.Q_LONGDIV, q_mul=1.0, q_div=1.0, gamma=gamma,
                                       direction=direction)
        Vo, Qo, _, _ = oracle.dense(spec)
        s, V, Q = _solve_all(S, spec)
        assert s.stats()["kernel_used"] == S.KERNEL_CASH_DIAG
        assert np.array_equal(V, Vo) and np.array_equal(Q, Qo), (price, v, K, means)


def test_integer_cash_kernel_is_used(S):
    for case in (cases.case_C_int, cases.case_C_int_K, cases.case_F_small):
        spec, _ = case()
        assert S.Solver(spec).solve().stats()["kernel_used"] == S.KERNEL_CASH_DIAG, spec.name
    # fractional cash grids: the cash-constraint kind shares its cash-independent terms across a row of cash levels,
    # the other kinds (whose balance depends on the demand through the cash level) stay on the generic kernel
    for case, used in ((cases.case_C_small, S.KERNEL_CASH_ROW), (cases.case_C_rich, S.KERNEL_CASH_ROW),
                       (cases.case_D_small, S.KERNEL_GENERIC)):
        spec, _ = case()
        assert S.Solver(spec).solve().stats()["kernel_used"] == used, spec.name


def test_cash_row_kernel_shapes(S, oracle):
    """bi_cash_row: several 128-level cash segments with a ragged last one, 0.1 and integer cash grids, K > 0 with a
    cash reserve, deposit and overhead rates, bankruptcy penalty, gamma < 1, survival recursion with per-period
    prices; whole grid against the oracle and the generic kernel."""
    specs = [
        S.cash_constraint_model(cases.pmf([4, 6, 5], 0.999), price=7, vari_cost=1.5, fixed_cost=3, hold_cost=0.25, salvage=0.5,
                                overhead=2, overhead_rate=0.05, deposit_rate=0.02, penalty_cost=0.3, max_order=11,
                                inv_min=0, inv_max=17, cash_min=0, cash_max=41.7, gamma=0.97),
        S.cash_constraint_model(cases.pmf([5, 5], 0.99), price=9, vari_cost=2, salvage=1, max_order=9, inv_min=0, inv_max=14,
                                cash_min=0, cash_max=300, quantiser=S.abi.Q_LONGDIV, q_mul=1.0, q_div=1.0, hold_cost=0.5),
        S.cash_survival_model(cases.pmf([5, 6, 4], 0.99), price_t=[4.5, 5, 4], vari_cost_t=[1, 1.5, 2], overhead_t=[11, 9, 12],
                              salvage=0.5, hold_cost=0.25, deposit_rate=0.01, max_order=12, inv_min=0, inv_max=20,
                              cash_min=-20, cash_max=150),
    ]
    for spec in specs:
        Vo, Qo, evals, _ = oracle.dense(spec)
        s, V, Q = _solve_all(S, spec)
        assert s.stats()["kernel_used"] == S.KERNEL_CASH_ROW and s.stats()["evals"] == evals, spec.name
        assert np.array_equal(V, Vo) and np.array_equal(Q, Qo), spec.name
        g, Vg, Qg = _solve_all(S, spec, kernel=S.KERNEL_GENERIC)
        assert g.stats()["kernel_used"] == S.KERNEL_GENERIC
        assert np.array_equal(Vg, Vo) and np.array_equal(Qg, Qo), spec.name


@pytest.mark.parametrize("case", [cases.case_B1_ref, cases.case_B1_fixed, cases.case_B2_small],
                         ids=lambda f: f.__name__[5:])
@pytest.mark.parametrize("dedup", [False, True])
def test_staged_kernel_whole_grid(case, dedup, S, oracle):
    """bi_backorder_staged requested explicitly (AUTO prefers the slab kernel)."""
    spec, _ = case()
    Vo, Qo, _, _ = oracle.dense(spec)
    s, V, Q = _solve_all(S, spec, kernel=S.KERNEL_STAGED, dedup=dedup)
    assert np.array_equal(V, Vo) and np.array_equal(Q, Qo)
    assert s.stats()["kernel_used"] == S.KERNEL_STAGED


def test_lead_slab_wide_cases(S, oracle):
    """Slab-kernel tile edges: preQ counts that are not multiples of 8 / 4, D not a multiple of 4,
    lead time 1 and 2, clamped and unclamped, with and without folding."""
    for (lead, max_order, means, clamp) in [(2, 9, [3, 4, 3], True), (2, 12, [5, 2], True), (1, 21, [6, 7, 5], True),
                                            (1, 10, [3, 3, 3], False), (2, 4, [2, 2, 2], False)]:
        p = S.poisson_pmf(means, 0.999)
        dmax = max(r[-1, 0] for r in p)
        T = len(means)
        inv_min, inv_max = (-T * dmax, T * max_order) if not clamp else (-14, 17)
        spec = S.leadtime_model(p, fixed_cost=4, vari_cost=1, hold_cost=2, penalty_cost=9, max_order=max_order,
                                inv_min=inv_min, inv_max=inv_max, lead_time=lead, clamp=clamp)
        Vo, Qo, _, _ = oracle.dense(spec)
        for dedup in (False, True):
            auto = S.KERNEL_LEAD_Q2 if lead == 2 and not dedup else S.KERNEL_LEAD_COL
            for k, used in ((S.KERNEL_AUTO, auto), (S.KERNEL_LEAD_COL, S.KERNEL_LEAD_COL),
                            (S.KERNEL_LEAD_SLAB, S.KERNEL_LEAD_SLAB)):
                s, V, Q = _solve_all(S, spec, dedup=dedup, kernel=k)
                assert s.stats()["kernel_used"] == used
                assert np.array_equal(V, Vo) and np.array_equal(Q, Qo), (lead, max_order, clamp, dedup, k)


def test_staged_kernel_is_used_for_leadtime(S):
    spec, _ = cases.case_B2_small()
    assert S.Solver(spec).solve().stats()["kernel_used"] == S.KERNEL_LEAD_Q2
    assert S.Solver(spec, kernel=S.KERNEL_LEAD_COL).solve().stats()["kernel_used"] == S.KERNEL_LEAD_COL
    assert S.Solver(spec, kernel=S.KERNEL_LEAD_SLAB).solve().stats()["kernel_used"] == S.KERNEL_LEAD_SLAB
    with pytest.raises(S.SdpbError):  # folded grids and lead time 1 have no all-actions-in-thread kernel
        S.Solver(spec, kernel=S.KERNEL_LEAD_Q2, dedup=True)
    with pytest.raises(S.SdpbError):
        S.Solver(cases.case_B1_ref()[0], kernel=S.KERNEL_LEAD_Q2)
    assert S.Solver(spec, kernel=S.KERNEL_STAGED).solve().stats()["kernel_used"] == S.KERNEL_STAGED
    assert S.Solver(spec, kernel=S.KERNEL_GENERIC).solve().stats()["kernel_used"] == S.KERNEL_GENERIC
    spec, _ = cases.case_A_small()
    assert S.Solver(spec).solve().stats()["kernel_used"] == S.KERNEL_FUSED  # small grid: the whole horizon in one launch
    assert S.Solver(spec, kernel=S.KERNEL_TILED).solve().stats()["kernel_used"] == S.KERNEL_TILED


@pytest.mark.parametrize("case", A_CASES, ids=lambda f: f.__name__[5:])
def test_fused_whole_horizon_kernel(case, S, oracle):
    """bi_inv_fused: one cooperative launch for all periods (small unsharded 1-D grids).  Whole grid against the
    oracle, repeated solves, and the per-period API (which keeps using the per-period kernels) on the same handle."""
    spec, init = case()
    Vo, Qo, evals, _ = oracle.dense(spec)
    if spec.flags & S.abi.F_GY_MODE:
        with pytest.raises(S.SdpbError):
            S.Solver(spec, kernel=S.KERNEL_FUSED)
        return
    s, V, Q = _solve_all(S, spec, kernel=S.KERNEL_FUSED)
    st = s.stats()
    assert st["kernel_used"] == S.KERNEL_FUSED and st["launches"] == 1 and st["evals"] == evals
    assert np.array_equal(V, Vo) and np.array_equal(Q, Qo)
    for _ in range(2):
        s.solve()
        V1, Q1 = s.period_tables(1)
        assert np.array_equal(V1, Vo[0]) and np.array_equal(Q1, Qo[0])
    for t in range(spec.T, 0, -1):
        s.solve_period_async(t)
    s.sync()
    V1, Q1 = s.period_tables(1)
    assert np.array_equal(V1, Vo[0]) and np.array_equal(Q1, Qo[0])


def test_fused_kernel_shapes(S, oracle):
    """States not a multiple of the CTA slice, more states than SMs (several states per CTA), fewer states than
    SMs, MAX direction, sparse demand support, action counts around the 8-pair batch and the 32-lane argopt."""
    rng = np.random.default_rng(3)
    for n_states, n_act, direction in ((149, 33, S.abi.MIN), (700, 9, S.abi.MAX), (2001, 40, S.abi.MIN), (37, 257, S.abi.MIN),
                                       (5000, 3, S.abi.MIN)):
        rows = []
        for _ in range(3):
            sup = np.sort(rng.choice(np.arange(0, 14), size=int(rng.integers(2, 9)), replace=False)).astype(float)
            rows.append(np.column_stack([sup, rng.dirichlet(np.ones(len(sup)))]))
        half = n_states // 2
        spec = S.inventory_model(rows, fixed_cost=7, vari_cost=1, hold_cost=1, penalty_cost=6, max_order=n_act - 1,
                                 inv_min=-half, inv_max=n_states - half - 1, direction=direction)
        Vo, Qo, evals, _ = oracle.dense(spec)
        s, V, Q = _solve_all(S, spec, kernel=S.KERNEL_FUSED)
        assert s.stats()["kernel_used"] == S.KERNEL_FUSED
        assert np.array_equal(V, Vo) and np.array_equal(Q, Qo), (n_states, n_act)
    with pytest.raises(S.SdpbError):
        S.Solver(cases.case_B1_ref()[0], kernel=S.KERNEL_FUSED)
    # AUTO: C1 (1,001 states) runs fused, C2 (2,001 x 101) stays on the per-period tiled kernel; same tables either way
    assert S.Solver(S.configs.c1()).solve().stats()["kernel_used"] == S.KERNEL_FUSED
    spec = S.configs.c2()
    a, Va, Qa = _solve_all(S, spec)
    f, Vf, Qf = _solve_all(S, spec, kernel=S.KERNEL_FUSED)
    assert a.stats()["kernel_used"] == S.KERNEL_TILED and f.stats()["kernel_used"] == S.KERNEL_FUSED
    assert np.array_equal(Va, Vf) and np.array_equal(Qa, Qf)


@pytest.mark.parametrize("name", cases.GOLDEN)
def test_against_golden_fixtures(name, S):
    spec, init, g = cases.load_golden(name)
    s, V, Q = _solve_all(S, spec)
    per = g["periods"] - 1
    assert np.array_equal(V[per], g["V"])
    assert np.array_equal(Q[per].astype(np.float32), g["Q"])
    v, q = s.value(1, g["init"])
    assert np.array_equal(v, g["init_values"])


@pytest.mark.parametrize("case", cases.ALL, ids=lambda f: f.__name__[5:])
def test_value_and_opt_table_match_topdown(case, S, oracle):
    """getExpectedValue/getAction and the getOptTable() row set (visited states only, sorted)."""
    spec, init = case()
    rows, iv, _ = oracle.topdown(spec, init)
    s = S.Solver(spec).solve()
    v, q = s.value(1, init)
    assert np.array_equal(v, iv)
    s.reach(init)
    tab = s.opt_table()
    nd = spec.ndim
    assert tab.shape == (len(rows), nd + 2)
    # the oracle emits rows in the memo-map order (period, inv, preQ.., cash); the library in its
    # grid order, which is the same ordering
    assert np.array_equal(tab[:, :nd + 1], rows[:, :nd + 1])
    assert np.array_equal(tab[:, -1], rows[:, -2])
    # every visited state answers through sdpb_value too
    for t in range(1, spec.T + 1):
        sel = rows[rows[:, 0] == t]
        if len(sel):
            v, q = s.value(t, sel[:, 1:1 + nd])
            assert np.array_equal(v, sel[:, -1]) and np.array_equal(q, sel[:, -2])


@pytest.mark.parametrize("case", cases.ALL, ids=lambda f: f.__name__[5:])
def test_device_lambdas_match_oracle(case, S, oracle):
    """sdpb_eval_triples (the device code of c, f, |A|) against the oracle's lambdas."""
    import random
    spec, _ = case()
    if spec.two_product or spec.staff:
        pytest.skip("sdpb_eval_triples is not implemented for the two-product and staff kinds")
    pkg = S.package
    from importlib import import_module
    spot = import_module(pkg.__name__ + "._spot")
    rng = random.Random(7)
    s = S.Solver(spec)
    for _ in range(60):
        st, a, d = spot.random_triple(spec, rng)
        c, nxt, na = spot.eval_descriptor(spec, st, a, d, solver=s)
        assert na == oracle.n_actions(spec, st[0], st[1:])
        co, no = oracle.eval_triple(spec, st[0], st[1:], a, d)
        assert c == co
        if st[0] < spec.T and oracle.index(spec, no) >= 0:
            assert tuple(no) == nxt


def test_repeated_solves_replay_a_cuda_graph(S, oracle):
    """From the second sdpb_solve on, the T launches are one captured graph: results must not change."""
    for case in (cases.case_A_small, cases.case_C_int, cases.case_B2_small):
        spec, init = case()
        Vo, Qo, evals, _ = oracle.dense(spec)
        s = S.Solver(spec)
        for rep in range(4):
            s.solve()
            for t in (1, spec.T):
                V, Q = s.period_tables(t)
                assert np.array_equal(V, Vo[t - 1]) and np.array_equal(Q, Qo[t - 1]), (spec.name, rep, t)
            assert s.stats()["evals"] == evals
        s.solve_async()
        s.sync()
        v, _ = s.value(1, init)
        assert v[0] == Vo[0][oracle.index(spec, init[0])]


@pytest.mark.parametrize("case", [cases.case_A_small, cases.case_B1_fixed, cases.case_B2_small, cases.case_C_rich,
                                  cases.case_C_int, cases.case_D_rich, cases.case_E_small, cases.case_XR_small,
                                  cases.case_DT_small, cases.case_TP_small], ids=lambda f: f.__name__[5:])
def test_policy_rollout_matches_oracle(case, S, oracle):
    """sdpb_simulate (Simulation.java:59-70) against the oracle's roll-out of its own policy tables,
    path by path, on seeded samples that include out-of-support demands and non-integer draws."""
    spec, init = case()
    Vo, Qo, _, _ = oracle.dense(spec)
    rng = np.random.default_rng(5)
    dmax = max(r[-1, 0] for r in spec.pmf)
    samples = rng.uniform(-0.4, dmax + 3.0, size=(500, spec.T))
    gamma = spec.gamma
    s = S.Solver(spec).solve()
    got = s.simulate(init[0], samples, gamma)
    want = oracle.simulate(spec, Qo, init[0], samples, gamma)
    assert np.array_equal(got, want)


def test_reference_style_simulation(S, oracle):
    """Reads like CLSPTesting.java:125-127: Simulation(distributions, sampleNum, recursion)."""
    dists = [S.PoissonDist(m) for m in (5, 8, 6)]
    pmf = S.GetPmf(dists, 0.999, 1).getpmf()
    spec = S.inventory_model(pmf, 20, 1, 1, 5, max_order=15, inv_min=-30, inv_max=30)
    recursion = S.Recursion(spec)
    sim = S.Simulation(dists, 2000, recursion)
    samples = S.generate_lh_samples(dists, 2000)
    mean = sim.simulateSDPGivenSamplNum(S.State(1, 0), samples)
    opt = recursion.getExpectedValue(S.State(1, 0))
    assert abs(mean - opt) / opt < 0.05          # the reference's eyeball check: simulated ~ optimal
    Vo, Qo, _, _ = oracle.dense(spec)
    assert np.array_equal(sim.simulate_paths(S.State(1, 0), samples), oracle.simulate(spec, Qo, [0.0], samples))


@pytest.mark.parametrize("case,world", [(cases.case_A_small, 2), (cases.case_A_gy, 3), (cases.case_B2_small, 3),
                                        (cases.case_B1_ref, 2), (cases.case_C_int, 3), (cases.case_C_rich, 2),
                                        (cases.case_F_small, 2), (cases.case_E_small, 2), (cases.case_M2_small, 3),
                                        (cases.case_W_small, 2)], ids=lambda x: getattr(x, "__name__", str(x))[5:])
def test_sharded_handles_on_one_gpu(case, world, S, oracle):
    """The multi-GPU data path on one device: `world` handles with shard_rank 0..world-1, stepped period
    by period; after each period every rank's block of V_t is copied into every other rank's table (what
    the NCCL all-gather does).  Checks the kernels' [lo, hi) handling bit for bit."""
    import torch
    par = S.package.parallel
    spec, _ = case()
    Vo, Qo, evals, _ = oracle.dense(spec)
    n = Vo.shape[1]
    hs = [S.Solver(spec, shard_rank=r, shard_count=world) for r in range(world)]
    bounds = [par.shard_bounds(n, r, world) for r in range(world)]
    chunk = bounds[0][2]
    for t in range(spec.T, 0, -1):
        for h in hs:
            h.solve_period_async(t)
        for h in hs:
            h.sync()
        tabs = [par.wrap_device(torch, h.device_tables(t)[0], chunk * world, "<f8", 0) for h in hs]
        for r, (lo, hi, _) in enumerate(bounds):
            for r2 in range(world):
                if r2 != r:
                    tabs[r2][lo:hi].copy_(tabs[r][lo:hi])
        torch.cuda.synchronize()
    total = 0.0
    for r, (lo, hi, _) in enumerate(bounds):
        for t in range(1, spec.T + 1):
            dv, dq = hs[r].device_tables(t)
            V = par.wrap_device(torch, dv, chunk * world, "<f8", 0)[:n].cpu().numpy()
            Qi = par.wrap_device(torch, dq, chunk * world, "<i4", 0)[lo:hi].cpu().numpy()
            assert np.array_equal(V, Vo[t - 1]), (r, t)
            # policy: action index -> order quantity (pair index for the two-product kind)
            q = np.where(Qi < 0, 0.0, Qi * spec.step)
            if spec.cost_kind == S.COST_CASH_XR:
                pytest.skip("XR policy needs the state's inventory")
            assert np.array_equal(q, Qo[t - 1][lo:hi]), (r, t)
        total += hs[r].stats()["evals"]
    assert total == evals


@pytest.mark.parametrize("case,world", [(cases.case_A_small, 3), (cases.case_A_gy, 3), (cases.case_B2_small, 4),
                                        (cases.case_B1_ref, 3), (cases.case_C_int, 3), (cases.case_C_rich, 2),
                                        (cases.case_D_small, 3), (cases.case_M2_small, 3), (cases.case_W_small, 2),
                                        (cases.case_XR_small, 2)], ids=lambda x: getattr(x, "__name__", str(x))[5:])
def test_halo_exchange_ranges_on_one_gpu(case, world, S, oracle):
    """sdpb_shard_reads: each rank is given ONLY the rows of V_{t+1} it declared (everything else in its table
    is NaN) and must still produce the oracle's block; the host mirror of the range agrees with the library."""
    import torch
    par = S.package.parallel
    spec, _ = case()
    Vo, Qo, _, _ = oracle.dense(spec)
    n = Vo.shape[1]
    hs = [S.Solver(spec, shard_rank=r, shard_count=world) for r in range(world)]
    bounds = [par.shard_bounds(n, r, world) for r in range(world)]
    chunk = bounds[0][2]
    needs = [h.shard_reads() for h in hs]
    for r, (lo, hi, _) in enumerate(bounds):
        assert needs[r] == par.needed_range(spec, lo, hi, n), (r, needs[r])
    for t in range(spec.T, 0, -1):
        tabs = [par.wrap_device(torch, h.device_tables(t)[0], chunk * world, "<f8", 0) for h in hs]
        for tab in tabs:
            tab.fill_(float("nan"))
        torch.cuda.synchronize()
        for h in hs:
            h.solve_period_async(t)
        for h in hs:
            h.sync()
        for r, (lo, hi, _) in enumerate(bounds):       # what HaloExchange sends: overlap(block[r], needs[r2])
            for r2 in range(world):
                a, b = max(lo, needs[r2][0]), min(hi, needs[r2][1])
                if r2 != r and b > a:
                    tabs[r2][a:b].copy_(tabs[r][a:b])
        torch.cuda.synchronize()
        for r, (lo, hi, _) in enumerate(bounds):
            assert np.array_equal(tabs[r][lo:hi].cpu().numpy(), Vo[t - 1][lo:hi]), (r, t)


def test_large_sharded_c5_on_one_gpu(S, oracle):
    """tiled2 with shard boundaries that cut through 1021-state tiles."""
    import torch
    par = S.package.parallel
    spec = S.configs.c5(n_states=300_007, T=2, n_actions=37)
    ref = S.Solver(spec).solve()
    world = 3
    hs = [S.Solver(spec, shard_rank=r, shard_count=world, kernel=S.KERNEL_TILED2) for r in range(world)]
    n = ref.n_states
    bounds = [par.shard_bounds(n, r, world) for r in range(world)]
    chunk = bounds[0][2]
    for t in (2, 1):
        for h in hs:
            h.solve_period_async(t)
        for h in hs:
            h.sync()
        tabs = [par.wrap_device(torch, h.device_tables(t)[0], chunk * world, "<f8", 0) for h in hs]
        for r, (lo, hi, _) in enumerate(bounds):
            for r2 in range(world):
                if r2 != r:
                    tabs[r2][lo:hi].copy_(tabs[r][lo:hi])
        torch.cuda.synchronize()
    assert hs[0].stats()["kernel_used"] == S.KERNEL_TILED2
    for t in (1, 2):
        Vr, Qr = ref.period_tables(t)
        for r, (lo, hi, _) in enumerate(bounds):
            dv, dq = hs[r].device_tables(t)
            V = par.wrap_device(torch, dv, chunk * world, "<f8", 0)[:n].cpu().numpy()
            Qi = par.wrap_device(torch, dq, chunk * world, "<i4", 0)[lo:hi].cpu().numpy()
            assert np.array_equal(V, Vr)
            assert np.array_equal(Qi * spec.step, Qr[lo:hi])


def test_unsolved_state_raises(S):
    spec, _ = cases.case_A_small()
    s = S.Solver(spec)
    with pytest.raises(S.SdpbError) as e:
        s.value(1, [[0.0]])
    assert e.value.code == S.abi.SDPB_ERR_STATE
    s.solve()
    with pytest.raises(S.SdpbError) as e:
        s.value(1, [[1000.0]])  # outside the grid: Java would throw from getAction
    assert e.value.code == S.abi.SDPB_ERR_UNSOLVED
    with pytest.raises(S.SdpbError):
        s.value(1, [[0.25]])    # not a grid point


def test_reference_style_driver_capacitated(S, oracle):
    """Reads like src/capacitated/CLSPTesting.java:75-120: build pmf, lambdas, Recursion, solve."""
    pmf = S.GetPmf([S.PoissonDist(m) for m in (5, 8, 6)], 0.999, 1).getpmf()
    K, v, h, pi, maxQ, lo, hi = 20.0, 1.0, 1.0, 5.0, 15, -30.0, 30.0

    def getFeasibleAction(s):
        return [float(i) for i in range(maxQ + 1)]

    def stateTransition(state, action, randomDemand):
        nxt = state.getIniInventory() + action - randomDemand
        nxt = hi if nxt > hi else nxt
        nxt = lo if nxt < lo else nxt
        return S.State(state.getPeriod() + 1, nxt)

    def immediateValue(state, action, randomDemand):
        fixed = K if action > 0 else 0
        lvl = state.getIniInventory() + action - randomDemand
        return fixed + v * action + h * max(lvl, 0) + pi * max(-lvl, 0)

    spec = S.inventory_model(pmf, K, v, h, pi, max_order=maxQ, inv_min=lo, inv_max=hi)
    recursion = S.Recursion(spec, getFeasibleAction, stateTransition, immediateValue)
    with pytest.raises(KeyError):
        recursion.getAction(S.State(1, 0))      # NullPointerException in the reference
    initialState = S.State(1, 0)
    finalValue = recursion.getExpectedValue(initialState)
    rows, iv, _ = oracle.topdown(spec, [[0.0]])
    assert finalValue == iv[0]
    assert recursion.getAction(initialState) == rows[0][-2]
    assert np.array_equal(recursion.getOptTable(), rows[:, :3])
    acts = recursion.getCacheActions()
    assert len(acts) == len(rows) and acts[S.State(1, 0)] == rows[0][-2]
    # a wrong descriptor is caught by the spot check
    wrong = S.inventory_model(pmf, K, v, h, pi + 1, max_order=maxQ, inv_min=lo, inv_max=hi)
    with pytest.raises(ValueError):
        S.Recursion(wrong, getFeasibleAction, stateTransition, immediateValue)


def test_reference_style_driver_cash(S, oracle):
    """Reads like src/cash/singleItem/CashConstraint.java:91-146."""
    spec, init = cases.case_C_rich()
    recursion = S.CashRecursion(spec)
    recursion.setTreeMapCacheAction()
    s0 = S.CashState(1, *init[0])
    rows, iv, _ = oracle.topdown(spec, init[:1])
    assert recursion.getExpectedValue(s0) == iv[0]
    assert recursion.getAction(s0) == rows[0][-2]
    assert np.array_equal(recursion.getOptTable(), rows[:, :4])


def test_reference_style_driver_two_product(S, oracle):
    """Reads like src/cash/multiItem/MultiItemCash.java:123-139 (commented-out solve block)."""
    spec, init = cases.case_M2_small()
    recursion = S.CashRecursionMulti(spec)
    iniState = S.CashStateMulti(1, *init[0])
    rows, iv, _ = oracle.topdown(spec, init)
    assert recursion.getExpectedValue(iniState) == iv[0]
    n = spec.max_order_idx + 1
    first = rows[0]
    assert recursion.getAction(iniState) == S.Actions(int(first[-2]) // n, int(first[-2]) % n)


def test_two_product_tolerance_changes_the_answer(S, oracle):
    """`> val + 0.1` is not a plain maximum: with the tolerance switched off the policy differs somewhere,
    and both variants match the oracle."""
    spec, _ = cases.case_M2_small()
    Va, Qa, _, _ = oracle.dense(spec)
    spec0, _ = cases.case_M2_small()
    spec0.tie_tolerance = 0.0
    Vb, Qb, _, _ = oracle.dense(spec0)
    assert not np.array_equal(Qa, Qb)
    for sp, Vo, Qo in ((spec, Va, Qa), (spec0, Vb, Qb)):
        s, V, Q = _solve_all(S, sp)
        assert np.array_equal(V, Vo) and np.array_equal(Q, Qo)


def test_two_product_row_kernel(S, oracle):
    """The row-shared two-product kernel (integer prices): several cash segments per (inv1, inv2) pair with a
    ragged last segment, gamma < 1, non-integer salvage (last period only), demand lists that are not product
    grids; whole grid against the oracle, the thread-per-state kernel and sharded handles.  Non-integer
    prices must fall back to the thread-per-state kernel."""
    rng = np.random.default_rng(11)
    for cash_max, qb, inv_max, price, vc, sal, gamma in ((300, 5, 6, (4, 9), (2, 4), (1, 1.5), 0.95),
                                                        (131, 7, 4, (6, 5), (3, 1), (0.5, 2), 1.0),
                                                        (127, 3, 5, (3, 8), (1, 2), (1, 1), 1.0)):
        rows = []
        for _ in range(3):
            n = int(rng.integers(3, 9))
            d = np.stack([rng.integers(0, 6, n), rng.integers(0, 5, n)], axis=1).astype(float)
            rows.append(np.column_stack([d, rng.dirichlet(np.ones(n))]))
        spec = S.two_product_cash_model(rows, price=price, vari_cost=vc, salvage=sal, q_bound=qb, inv_max=inv_max,
                                        cash_min=0, cash_max=cash_max, gamma=gamma)
        Vo, Qo, evals, _ = oracle.dense(spec)
        s, V, Q = _solve_all(S, spec)
        assert s.stats()["kernel_used"] == S.KERNEL_TWO_PRODUCT_ROW and s.stats()["evals"] == evals
        assert np.array_equal(V, Vo) and np.array_equal(Q, Qo)
        g, Vg, Qg = _solve_all(S, spec, kernel=S.KERNEL_GENERIC)
        assert g.stats()["kernel_used"] == S.KERNEL_GENERIC
        assert np.array_equal(Vg, Vo) and np.array_equal(Qg, Qo)
    spec = S.two_product_cash_model(rows, price=(4.5, 9), vari_cost=(2, 4), salvage=(1, 1), q_bound=4, inv_max=5,
                                    cash_min=0, cash_max=60)
    Vo, Qo, _, _ = oracle.dense(spec)
    s, V, Q = _solve_all(S, spec)
    assert s.stats()["kernel_used"] == S.KERNEL_GENERIC
    assert np.array_equal(V, Vo) and np.array_equal(Q, Qo)


def test_reference_style_driver_workforce(S, oracle):
    """Reads like src/workforce/WorkforcePlanning.java:106-118."""
    spec, init = cases.case_W_small()
    recursion = S.StaffRecursion(spec)
    initialState = S.StaffState(1, 0)
    rows, iv, _ = oracle.topdown(spec, init)
    assert recursion.getExpectedValue(initialState) == iv[0]
    assert recursion.getAction(initialState) == int(rows[0][-2])


def test_reference_style_driver_survival(S, oracle):
    spec, init = cases.case_F_small()
    recursion = S.RiskRecursion(spec)
    s0 = S.RiskState(1, init[0][0], init[0][1], False)
    rows, iv, _ = oracle.topdown(spec, init)
    assert recursion.getSurvProb(s0) == iv[0]
    assert 0.0 <= iv[0] <= 1.0


# ---- full-size configurations ----------------------------------------------------------------
def _sample_check(S, oracle, spec, s, n=48, seed=3):
    """Period-by-period inductive check on a sample: recompute V_t, Q_t on the CPU from the GPU's
    own V_{t+1} table and compare bits."""
    rng = np.random.default_rng(seed)
    idx = np.unique(np.concatenate([[0, s.n_states - 1], rng.integers(0, s.n_states, n)]))
    Vn = None
    for t in range(spec.T, 0, -1):
        V, Q = s.period_tables(t)
        vo, qo = oracle.step_states(spec, t, Vn, idx)
        assert np.array_equal(V[idx], vo), f"period {t}"
        assert np.array_equal(Q[idx], qo), f"period {t}"
        Vn = V


def test_config_c1_full(S, oracle):
    spec = S.configs.c1()
    Vo, Qo, evals, _ = oracle.dense(spec)
    for kernel in (S.KERNEL_GENERIC, S.KERNEL_AUTO):
        s, V, Q = _solve_all(S, spec, kernel=kernel)
        assert np.array_equal(V, Vo) and np.array_equal(Q, Qo)
        assert s.stats()["evals"] == evals == 133399266.0


def test_config_c2_full(S, oracle):
    spec = S.configs.c2()
    Vo, Qo, evals, _ = oracle.dense(spec)
    s, V, Q = _solve_all(S, spec)
    assert np.array_equal(V, Vo) and np.array_equal(Q, Qo)


def test_config_c3_sampled(S, oracle):
    spec = S.configs.c3(T=3)          # full 501 x 2001 grid, 3 of the 12 periods
    s = S.Solver(spec).solve()
    assert s.n_states == 1002501 and s.stats()["kernel_used"] == S.KERNEL_CASH_DIAG
    _sample_check(S, oracle, spec, s, n=24)
    for k in (S.KERNEL_GENERIC, S.KERNEL_CASH_INT):
        g = S.Solver(spec, kernel=k).solve()
        for t in (1, 2, 3):
            Va, Qa = s.period_tables(t)
            Vg, Qg = g.period_tables(t)
            assert np.array_equal(Va, Vg) and np.array_equal(Qa, Qg)


def test_config_c4_sampled(S, oracle):
    spec = S.configs.c4(T=3)          # full 1001 x 101 x 101 grid, 3 of the 20 periods
    s = S.Solver(spec).solve()
    assert s.n_states == 10211201
    _sample_check(S, oracle, spec, s, n=64)
    # V(x, q1, q2) depends on (x, q1) only through x + q1: an exact structural identity
    V, Q = s.period_tables(1)
    V3 = V.reshape(1001, 101, 101)
    assert np.array_equal(V3[100, 7, :], V3[107, 0, :]) and np.array_equal(V3[500, 50, :], V3[520, 30, :])
    # the folded solve and the generic kernel give the same tables as the staged brute-force kernel
    assert s.stats()["kernel_used"] == S.KERNEL_LEAD_Q2
    for kw in ({"dedup": True}, {"kernel": S.KERNEL_GENERIC}, {"kernel": S.KERNEL_STAGED}, {"kernel": S.KERNEL_LEAD_SLAB},
               {"kernel": S.KERNEL_LEAD_COL}):
        d = S.Solver(spec, **kw).solve()
        Vd, Qd = d.period_tables(1)
        assert np.array_equal(Vd, V) and np.array_equal(Qd, Q)


def test_config_c5_sampled(S, oracle):
    spec = S.configs.c5(n_states=1_000_000, T=2)
    s = S.Solver(spec).solve()
    _sample_check(S, oracle, spec, s, n=64)
    assert s.stats()["kernel_used"] == S.KERNEL_TILED2
    # the tiled variants and the generic kernel are independent implementations: whole-grid agreement
    for k in (S.KERNEL_GENERIC, S.KERNEL_TILED):
        g = S.Solver(spec, kernel=k).solve()
        for t in (1, 2):
            Va, Qa = s.period_tables(t)
            Vg, Qg = g.period_tables(t)
            assert np.array_equal(Va, Vg) and np.array_equal(Qa, Qg)
