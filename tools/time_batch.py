"""sdpb_solve_batch on n copies of C2 (and single C1 / C2 latency)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpb200 as S
peak = S.abi.microbench(0)["nofma_tops"]
kern = {"tiled": S.KERNEL_TILED, "tiled2": S.KERNEL_TILED2, "auto": S.KERNEL_AUTO}[os.environ.get("BATCH_KERNEL", "auto")]
for n in (16, 64, 256):
    sp = S.configs.c2()
    batch = [S.Solver(sp, device=0, kernel=kern) for _ in range(n)]
    for _ in range(3):
        S.solve_batch(batch)
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps):
        S.solve_batch(batch)
    dt = (time.perf_counter() - t0) / reps
    st = batch[0].stats()
    print(f"batch {n:4d}: {dt*1e3:8.3f} ms  {dt*1e3/n:.4f} ms/instance  fp64 frac {st['fp64_ops']*n/dt/1e12/peak:.3f}", flush=True)
    for b in batch: b.close()
if os.environ.get("BATCH_ONLY"):
    sys.exit(0)
for name, mk in (("c1", S.configs.c1), ("c2", S.configs.c2)):
    for kern, kn in ((S.KERNEL_AUTO, "auto"), (S.KERNEL_TILED, "tiled"), (S.KERNEL_TILED2, "tiled2"), (S.KERNEL_FUSED, "fused")):
        try:
            s = S.Solver(mk(), device=0, kernel=kern)
        except S.SdpbError as e:
            print(name, kn, "n/a"); continue
        for _ in range(3): s.solve_async(); s.sync()
        t0 = time.perf_counter(); reps = 50
        for _ in range(reps): s.solve_async()
        s.sync()
        dt = (time.perf_counter() - t0) / reps
        print(f"{name} {kn:6s}: {dt*1e3:.4f} ms kernel_used={s.stats()['kernel_used']}", flush=True)
        s.close()
