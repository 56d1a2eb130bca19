"""world_size-2 (and 3) test of the multi-GPU schedule on CPU with the gloo backend: the block
partition, the padded all-gather of V_t every period and the period ordering are the code the GPU
path runs (stochastic-inventory_b200/parallel.py); the per-block solve is played by the CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case_name, out_dir, mode):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cases
    import oracle_lib as O
    import sdpb200 as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec, _ = getattr(cases, case_name)()
    n, _ = O.grid(spec)
    par = S.package.parallel
    lo, hi, chunk = par.shard_bounds(n, rank, world)
    full = [torch.full((chunk * world,), float("nan"), dtype=torch.float64) for _ in range(spec.T)]
    Q = np.full((spec.T, n), np.nan)

    def solve_block(t):
        vn = None if t == spec.T else full[t].numpy()[:n]
        idx = np.arange(lo, hi, dtype=np.int64)
        v, q = O.step_states(spec, t, vn, idx)
        full[t - 1][lo:hi] = torch.from_numpy(v)
        Q[t - 1, lo:hi] = q

    exchange, need = None, (0, n)
    if mode == "halo":
        need = par.needed_range(spec, lo, hi, n)
        needs = par.gather_needs(torch, dist, world, need)
        assert needs == [par.needed_range(spec, *par.shard_bounds(n, r, world)[:2], n) for r in range(world)]
        exchange = par.HaloExchange(dist, rank, world, n, needs)
    par.backward_induction_sharded(spec.T, n, rank, world, solve_block, full, dist.all_gather_into_tensor, exchange)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), V=torch.stack(full).numpy()[:, :n], Q=Q, lo=lo, hi=hi,
             need=np.array(need))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,case_name,mode", [(2, "case_A_small", "allgather"), (2, "case_C_rich", "allgather"),
                                                  (3, "case_B2_small", "allgather"), (3, "case_A_gy", "halo"),
                                                  (3, "case_B2_small", "halo"), (2, "case_C_int", "halo")])
def test_sharded_schedule_matches_unsharded(world, case_name, mode, tmp_path, oracle):
    import cases
    port = 29500 + (os.getpid() % 2000) + world + (7 if mode == "halo" else 0)
    mp.start_processes(_worker, args=(world, port, case_name, str(tmp_path), mode), nprocs=world, join=True,
                       start_method="spawn")
    spec, _ = getattr(cases, case_name)()
    Vo, Qo, _, _ = oracle.dense(spec)
    covered = np.zeros(Vo.shape[1], dtype=bool)
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        lo, hi = int(g["lo"]), int(g["hi"])
        a, b = (int(x) for x in g["need"])
        # every rank ends with the part of every V_t (t >= 2) it reads -- all of it under the all-gather -- and
        # with its own block of V_1, which nobody else reads
        assert np.array_equal(g["V"][1:, a:b], Vo[1:, a:b])
        assert np.array_equal(g["V"][:, lo:hi], Vo[:, lo:hi])
        if mode == "halo" and (a > 0 or b < Vo.shape[1]):
            assert np.isnan(g["V"][1:, :a]).all() and np.isnan(g["V"][1:, b:]).all()   # nothing else was sent
        assert np.array_equal(g["Q"][:, lo:hi], Qo[:, lo:hi])
        covered[lo:hi] = True
    assert covered.all()


def test_shard_bounds_cover_and_pad(S):
    par = S.package.parallel
    for n in (1, 7, 61, 1000, 10211201):
        for world in (1, 2, 3, 4, 8):
            pieces = [par.shard_bounds(n, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            for a, b in zip(pieces, pieces[1:]):
                assert a[1] == b[0]
            chunk = pieces[0][2]
            assert chunk * world >= n and all(p[1] - p[0] <= chunk for p in pieces)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_needed_range_is_sufficient_for_every_case(world, S, oracle):
    """The host mirror of sdpb_shard_reads, for every case of the suite: a block solved from a V_{t+1} that is NaN
    outside its declared range equals the unsharded solve (no process group needed: the oracle plays the kernel)."""
    import cases
    import oracle_lib as O
    par = S.package.parallel
    for case in cases.ALL:
        spec, _ = case()
        Vo, Qo, _, _ = oracle.dense(spec)
        n = Vo.shape[1]
        for r in range(world):
            lo, hi, _ = par.shard_bounds(n, r, world)
            if hi <= lo:
                assert par.needed_range(spec, lo, hi, n) == (0, 0)
                continue
            a, b = par.needed_range(spec, lo, hi, n)
            assert 0 <= a < b <= n
            idx = np.arange(lo, hi, dtype=np.int64)
            for t in range(spec.T - 1, 0, -1):   # periods that read a successor table
                vn = np.full(n, np.nan)
                vn[a:b] = Vo[t][a:b]
                v, q = O.step_states(spec, t, vn, idx)
                assert np.array_equal(v, Vo[t - 1][lo:hi]) and np.array_equal(q, Qo[t - 1][lo:hi]), (spec.name, r, t)
