"""Solve one configuration once (for ncu): python tools/run_one.py c3|c4|c5|c2|c1 [T] [shard_rank shard_count]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpb200 as S  # noqa: E402

name = sys.argv[1]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mk = {"c1": lambda: S.configs.c1(), "c2": lambda: S.configs.c2(), "c3": lambda: S.configs.c3(T=T),
      "c4": lambda: S.configs.c4(T=T), "c5": lambda: S.configs.c5(n_states=10_000_000, T=T)}[name]
spec = mk()
kw = {}
if len(sys.argv) > 4:
    kw = {"shard_rank": int(sys.argv[3]), "shard_count": int(sys.argv[4])}
s = S.Solver(spec, **kw)
for t in range(spec.T, 0, -1):
    s.solve_period_async(t)
s.sync()
print(name, s.stats())
