"""Solve time of the cash kinds on fractional / integer grids, AUTO against bi_generic (second solve timed)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpb200 as S

pm = S.poisson_pmf
jobs = {
    "deposit 0.1-grid (CashConstraint.java)": S.cash_constraint_model(pm([60.0] * 6, 0.9999), price=8, vari_cost=1.5, fixed_cost=10, hold_cost=0.25,
                                               salvage=0.5, overhead=5, overhead_rate=0.02, deposit_rate=0.01, penalty_cost=0.3,
                                               max_order=100, inv_min=0, inv_max=200, cash_min=0, cash_max=400, gamma=0.98),
    "overdraft int-grid 800k (CashOverdraft.java)": S.cash_overdraft_model(pm([60.0] * 4, 0.9999), price=10, vari_cost=1, overhead_t=[50] * 4, od_limit=500,
                                                    r0=0.01, r2=0.1, r3=2.0, max_order=100, inv_min=0, inv_max=200,
                                                    cash_min=-1000, cash_max=3000),
    "overdraft 0.1-grid 800k": S.cash_overdraft_model(pm([60.0] * 4, 0.9999), price=10, vari_cost=1, overhead_t=[50] * 4, od_limit=100,
                                                     r0=0.01, r2=0.1, r3=2.0, max_order=100, inv_min=0, inv_max=200,
                                                     cash_min=-100, cash_max=300, quantiser=S.Q_DIV, q_mul=10.0, q_div=10.0),
    "overdraft limit (CashOverdraftLimit.java)": S.cash_overdraft_limit_model(pm([60.0] * 4, 0.9999), price=10, vari_cost=1, hold_cost=0.5, overhead_t=[50] * 4,
                                                      max_order=100, inv_min=0, inv_max=200, cash_min=-1000, cash_max=3000),
}
only = sys.argv[1:] or None
for name, spec in jobs.items():
    if only and not any(o in name for o in only):
        continue
    for kernel, kname in ((S.KERNEL_AUTO, "auto"), (S.KERNEL_GENERIC, "generic")):
        s = S.Solver(spec, device=0, kernel=kernel)
        s.solve()
        t0 = time.perf_counter()
        s.solve()
        dt = time.perf_counter() - t0
        st = s.stats()
        print(f"{name:46s} {kname:8s} kernel_used={st['kernel_used']:2d} states={s.n_states} solve={dt*1e3:8.2f} ms "
              f"evals/s={st['evals']/dt:.3e}", flush=True)
        s.close()
