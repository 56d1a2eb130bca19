python -m pytest tests -m gpu -x -q -k "collapsed" 2>&1 | tail -15
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import sdpb200 as S, numpy as np
for n in (1_000_000, 10_000_000):
    sp = S.configs.c5(n_states=n)
    with S.Solver(sp, kernel=S.KERNEL_COLLAPSED) as co:
        co.solve(); co.solve()
        t0=time.perf_counter()
        for _ in range(5): co.solve_async()
        co.sync(); dt=(time.perf_counter()-t0)/5
        print(n, "collapsed ms", dt*1e3, co.stats()["evals_executed"]/co.stats()["evals"])
PY
