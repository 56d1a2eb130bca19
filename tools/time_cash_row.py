"""bi_cash_row against bi_generic on a CashConstraint.java-style instance with the reference's default 0.1 cash grid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpb200 as S

pmf = S.poisson_pmf([60.0] * 6, 0.9999)
spec = S.cash_constraint_model(pmf, price=8, vari_cost=1.5, fixed_cost=10, hold_cost=0.25, salvage=0.5, overhead=5,
                               overhead_rate=0.02, deposit_rate=0.01, penalty_cost=0.3, max_order=100, inv_min=0,
                               inv_max=200, cash_min=0, cash_max=400, gamma=0.98)   # quantiser round(w*10)/10.0
for kernel, name in ((S.KERNEL_AUTO, "auto"), (S.KERNEL_GENERIC, "generic")):
    s = S.Solver(spec, device=0, kernel=kernel)
    s.solve()
    t0 = time.perf_counter()
    s.solve()
    dt = time.perf_counter() - t0
    st = s.stats()
    v, q = s.value(1, [[0.0, 50.0]])
    print(f"{name:8s} kernel_used={st['kernel_used']:2d} states={s.n_states} demands={len(pmf[0])} "
          f"solve={dt*1e3:8.2f} ms  evals/s={st['evals']/dt:.3e}  V1={v[0]!r} Q1={q[0]}")
    s.close()
