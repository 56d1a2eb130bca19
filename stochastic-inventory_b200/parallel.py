"""Multi-GPU partitioning of one solve: the flattened state grid is cut into `world` contiguous
blocks (inventory outermost, so each rank owns a band of inventory levels); every period each rank
solves its block, then V_t is all-gathered so that every rank holds the full table period t-1
reads.  Q_t is never exchanged.  One process per GPU; the collective is NCCL over NVLink through
torch.distributed (gloo in the CPU tests).  SURVEY.md §8(e).
"""
from __future__ import annotations


def shard_bounds(n_states: int, rank: int, world: int):
    """-> (lo, hi, chunk): rank's block [lo, hi) and the padded block length the all-gather uses
    (same arithmetic as sdpb_create: chunk = ceil(S / world))."""
    chunk = (n_states + world - 1) // world
    lo = min(n_states, chunk * rank)
    hi = min(n_states, lo + chunk)
    return lo, hi, chunk


def backward_induction_sharded(T, n_states, rank, world, solve_block, full_tables, all_gather):
    """The per-period schedule shared by the GPU path and the CPU (gloo) test.

    solve_block(t)            computes V_t on [lo, hi) into full_tables[t-1][lo:hi], reading
                              full_tables[t] (all of it) when t < T
    full_tables[t-1]          a tensor of chunk*world elements (the padded full V_t)
    all_gather(out, inp)      torch.distributed.all_gather_into_tensor or an equivalent
    """
    lo, hi, chunk = shard_bounds(n_states, rank, world)
    for t in range(T, 0, -1):
        solve_block(t)
        if world > 1:
            full = full_tables[t - 1]
            all_gather(full, full[chunk * rank: chunk * (rank + 1)])


class _CAI:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def wrap_device(torch, ptr, n, typestr, device):
    return torch.as_tensor(_CAI(ptr, n, typestr), device=f"cuda:{device}")


class ShardedSolve:
    """One rank's share of a GPU solve (libsdpb200 handle created with shard_rank / shard_count)."""

    def __init__(self, Solver, torch, dist, spec, rank, world, device, stream, kernel=0, dedup=False):
        self.torch, self.dist, self.world, self.rank = torch, dist, world, rank
        self.solver = Solver(spec, device=device, shard_rank=rank, shard_count=world, kernel=kernel,
                             dedup=dedup, stream=stream.cuda_stream)
        g = self.solver.grid
        self.T, self.n = g.T, g.n_states
        self.lo, self.hi, self.chunk = shard_bounds(self.n, rank, world)
        assert (self.lo, self.hi) == (g.shard_lo, g.shard_hi)
        self.V = []
        if world > 1:
            for t in range(1, self.T + 1):
                dv, _ = self.solver.device_tables(t)
                self.V.append(wrap_device(torch, dv, self.chunk * world, "<f8", device))

    def step(self):
        if self.world == 1:
            self.solver.solve_async()  # one CUDA graph per solve after the first
            return
        backward_induction_sharded(
            self.T, self.n, self.rank, self.world, self.solver.solve_period_async, self.V,
            self.dist.all_gather_into_tensor if self.world > 1 else None)

    def close(self):
        self.solver.close()
