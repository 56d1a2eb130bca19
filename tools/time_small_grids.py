"""Times the per-period tiled path against the fused whole-horizon kernel on small 1-D grids (used to pick the
AUTO rule in sdpb_create)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpb200 as S

def timed(spec, kernel, reps=20):
    s = S.Solver(spec, device=0, kernel=kernel)
    for _ in range(3):
        s.solve_async()
    s.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        s.solve_async()
    s.sync()
    dt = (time.perf_counter() - t0) / reps
    s.close()
    return dt * 1e3

print("states actions D T   tiled_ms fused_ms")
for n_states in (501, 1001, 2001, 4001, 8001):
    for n_act in (26, 51, 101, 201, 501):
        for mean, T in ((20.0, 8),):
            pmf = S.poisson_pmf([mean] * T, 0.9999)
            half = n_states // 2
            spec = S.inventory_model(pmf, fixed_cost=100, vari_cost=0, hold_cost=1, penalty_cost=10, max_order=n_act - 1,
                                     inv_min=-half, inv_max=n_states - half - 1)
            a, b = timed(spec, S.KERNEL_TILED), timed(spec, S.KERNEL_FUSED)
            print(f"{n_states:6d} {n_act:5d} {len(pmf[0]):3d} {T:2d}   {a:8.4f} {b:8.4f}  {'fused' if b < a else 'tiled'}")
