import time, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import sdpb200 as S
sp = S.configs.c5(n_states=10_000_000)
keep = S.Solver(sp, device=0); keep.solve()
for i in range(4):
    t0 = time.perf_counter(); s = S.Solver(sp, device=0); t1 = time.perf_counter(); s.solve(); t2 = time.perf_counter(); s.close(); t3 = time.perf_counter()
    print(f"cycle {i}: create {1e3*(t1-t0):.1f} ms solve {1e3*(t2-t1):.1f} destroy {1e3*(t3-t2):.2f}", file=sys.stderr)
