"""Per-shard kernel time of a partitioned solve, measured on ONE GPU: a handle created with shard_rank r of n is
stepped period by period without any exchange (the successor tables hold whatever is in memory -- the kernels' run time
does not depend on the values), so the number is the pure compute share of rank r at N = n.  Comparing
max_r(time) * n with the unsharded time separates tail / wave effects from exchange cost in a multi-GPU run.

    python tools/time_shards.py c4 8        # config, world
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpb200 as S  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    worlds = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8]
    spec = {"c3": S.configs.c3, "c4": S.configs.c4, "c5": lambda: S.configs.c5(n_states=10_000_000)}[name]()
    for world in worlds:
        ranks = range(world) if world <= 4 else (0, 1, world // 2, world - 1)
        out = []
        for r in ranks:
            s = S.Solver(spec, shard_rank=r, shard_count=world) if world > 1 else S.Solver(spec)
            best = 1e9
            for rep in range(3):
                s.sync()
                t0 = time.perf_counter()
                for t in range(spec.T, 0, -1):
                    s.solve_period_async(t)
                s.sync()
                best = min(best, (time.perf_counter() - t0) * 1e3)
            out.append((r, best, s.grid.device_bytes / 1e6))
            s.close()
        worst = max(b for _, b, _ in out)
        print(f"{name} world={world}: " + "  ".join(f"r{r}: {b:.2f} ms ({mb:.0f} MB)" for r, b, mb in out) +
              f"   -> slowest x world = {worst * world:.1f} ms", flush=True)


if __name__ == "__main__":
    main()
