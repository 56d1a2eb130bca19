"""Multi-GPU partitioning of one solve: the flattened state grid is cut into `world` contiguous
blocks (inventory outermost, so each rank owns a band of inventory levels); every period each rank
solves its block, then V_t is exchanged so that every rank holds the part of the table period t-1
reads: either all of it (all-gather) or only the band of inventory rows its own block can reach
(`HaloExchange`: point-to-point sends of the overlaps, 4-7x fewer bytes on the lead-time grid at 8
GPUs).  Q_t is never exchanged.  One process per GPU; NCCL over NVLink through torch.distributed
(gloo in the CPU tests).  SURVEY.md §8(e).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_states: int, rank: int, world: int):
    """-> (lo, hi, chunk): rank's block [lo, hi) and the padded block length the all-gather uses
    (same arithmetic as sdpb_create: chunk = ceil(S / world))."""
    chunk = (n_states + world - 1) // world
    lo = min(n_states, chunk * rank)
    hi = min(n_states, lo + chunk)
    return lo, hi, chunk


def needed_range(spec, lo, hi, n_states):
    """Host mirror of sdpb_shard_reads: the contiguous range of flattened V_{t+1} indices the block [lo, hi)
    may read -- its band of inventory rows widened by the largest order / pipeline quantity upwards and the
    largest demand downwards, plus the zero row when lost-sales clamps lift negative levels to it."""
    from . import _abi as A
    if hi <= lo:
        return 0, 0
    if spec.staff:
        return 0, n_states
    n_inv = int(round((spec.inv_max - spec.inv_min) / spec.step)) + 1
    per_x = n_states // n_inv
    r0, r1 = lo // per_x, (hi - 1) // per_x
    dmin = dmax = 0
    for row in spec.pmf:
        d = np.asarray(row, dtype=np.float64)[:, 0]
        di = np.trunc(d) if spec.two_product else np.rint(d / spec.step)
        dmin, dmax = min(dmin, int(di.min())), max(dmax, int(di.max()))
    r0 -= dmax
    r1 += spec.max_order_idx - dmin
    if spec.flags & A.F_LOST_SALES:
        i_zero = int(math_round_half_up((0.0 - spec.inv_min) / spec.step))
        r0, r1 = min(r0, i_zero), max(r1, i_zero)
    r0, r1 = max(r0, 0), min(r1, n_inv - 1)
    return r0 * per_x, (r1 + 1) * per_x


def math_round_half_up(x):
    """Java Math.round."""
    import math
    return math.floor(x + 0.5)


class HaloExchange:
    """Point-to-point exchange of V_t over torch.distributed: every rank receives, from each peer, the overlap of the
    peer's block with the range it will read (`needs[rank]`), and sends the mirror image.  `needs` must be identical
    on all ranks (gather_needs).  Works with NCCL (grouped sends/receives on the current stream) and gloo.  `base`:
    flattened index of element 0 of the tensors passed to __call__ (a shard holds only a window of V_t).
    This is the host-side statement of what libsdpb200's own peer exchange does (sdpb_peer_attach); the GPU path
    uses the library's, the gloo tests use this one."""

    def __init__(self, dist, rank, world, n_states, needs, base=0):
        self.dist, self.rank, self.base = dist, rank, base
        blocks = [shard_bounds(n_states, r, world)[:2] for r in range(world)]
        self.recv = []  # (peer, lo, hi): slices of peers' blocks that I read
        self.send = []  # (peer, lo, hi): slices of my block that peers read
        for peer in range(world):
            if peer == rank:
                continue
            a, b = max(blocks[peer][0], needs[rank][0]), min(blocks[peer][1], needs[rank][1])
            if b > a:
                self.recv.append((peer, a, b))
            a, b = max(blocks[rank][0], needs[peer][0]), min(blocks[rank][1], needs[peer][1])
            if b > a:
                self.send.append((peer, a, b))
        self.bytes_in = sum(b - a for _, a, b in self.recv) * 8

    def __call__(self, full):
        dist, o = self.dist, self.base
        ops = [dist.P2POp(dist.isend, full[a - o:b - o], peer) for peer, a, b in self.send]
        ops += [dist.P2POp(dist.irecv, full[a - o:b - o], peer) for peer, a, b in self.recv]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


def gather_needs(torch, dist, world, need, device=None):
    """All ranks' (lo, hi) read ranges, identical everywhere (one tiny all-gather at set-up)."""
    mine = torch.tensor(list(need), dtype=torch.int64, device=device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [(int(t[0]), int(t[1])) for t in out]


def backward_induction_sharded(T, n_states, rank, world, solve_block, full_tables, all_gather, exchange=None):
    """The per-period schedule shared by the GPU path and the CPU (gloo) test.

    solve_block(t)            computes V_t on [lo, hi) into full_tables[t-1][lo:hi], reading
                              full_tables[t] (the rows its block can reach) when t < T
    full_tables[t-1]          a tensor of chunk*world elements (the padded full V_t)
    all_gather(out, inp)      torch.distributed.all_gather_into_tensor or an equivalent
    exchange(full)            optional: a HaloExchange used instead of the all-gather
    """
    lo, hi, chunk = shard_bounds(n_states, rank, world)
    for t in range(T, 0, -1):
        solve_block(t)
        if world > 1 and t > 1:  # V_1 is read by nobody
            full = full_tables[t - 1]
            if exchange is not None:
                exchange(full)
            else:
                all_gather(full, full[chunk * rank: chunk * (rank + 1)])


class _CAI:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def wrap_device(torch, ptr, n, typestr, device):
    return torch.as_tensor(_CAI(ptr, n, typestr), device=f"cuda:{device}")


def gather_blobs(torch, dist, world, blob: bytes, device=None):
    """Every shard's sdpb_peer_export() bytes, in rank order (one small all-gather at set-up)."""
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).clone()
    if device is not None:
        mine = mine.to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


class ShardedSolve:
    """One rank's share of a GPU solve: a libsdpb200 handle created with shard_rank / shard_count.

    exchange = "p2p" (default): the shards are connected once through CUDA IPC (sdpb_peer_export -> all-gather of
        the 256-byte blobs -> sdpb_peer_attach); a step is then ONE library call, sdpb_solve_async, and the rows of
        V_t travel between the GPUs' tables inside the library (peer copies + device-side flags).
    exchange = "nccl": the round-1 path, kept for comparison -- the library is stepped period by period and V_t is
        exchanged with torch.distributed (point-to-point halo, or an all-gather when a shard holds the whole table)."""

    def __init__(self, Solver, torch, dist, spec, rank, world, device, stream, kernel=0, dedup=False, exchange="p2p",
                 profile=False):
        self.torch, self.dist, self.world, self.rank = torch, dist, world, rank
        self.solver = Solver(spec, device=device, shard_rank=rank, shard_count=world, kernel=kernel,
                             dedup=dedup, stream=stream.cuda_stream, profile=profile)
        g = self.solver.grid
        self.T, self.n = g.T, g.n_states
        self.lo, self.hi, self.chunk = shard_bounds(self.n, rank, world)
        assert (self.lo, self.hi) == (g.shard_lo, g.shard_hi)
        self.V = []
        self.exchange, self.exchange_kind = None, "none"
        self.bytes_out = self.bytes_in = 0
        if world > 1 and exchange == "p2p":
            blobs = gather_blobs(torch, dist, world, self.solver.peer_export(), device=f"cuda:{device}")
            ok, why = 1, ""
            try:
                self.solver.peer_attach(blobs)
            except Exception as e:  # e.g. no peer-to-peer path between two of the GPUs
                ok, why = 0, str(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{device}")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag[0]) == 1:
                self.bytes_out, self.bytes_in = self.solver.peer_traffic()
                self.exchange_kind = "p2p"
            else:
                # every rank falls back together: a fresh handle, stepped period by period with torch.distributed
                if rank == 0:
                    import sys
                    print(f"[sdpb200] peer exchange unavailable ({why or 'a peer could not attach'}): using the "
                          "torch.distributed exchange", file=sys.stderr)
                self.solver.close()
                self.solver = Solver(spec, device=device, shard_rank=rank, shard_count=world, kernel=kernel,
                                     dedup=dedup, stream=stream.cuda_stream)
                g = self.solver.grid
                exchange = "nccl"
        if world > 1 and exchange != "p2p":
            wlo, whi = g.window_lo, g.window_hi
            for t in range(1, self.T + 1):
                dv, _ = self.solver.device_tables(t)
                self.V.append(wrap_device(torch, dv, whi - wlo, "<f8", device))
            needs = gather_needs(torch, dist, world, self.solver.shard_reads(), device=f"cuda:{device}")
            self.exchange_kind = "nccl-allgather"
            if not (wlo == 0 and whi == self.chunk * world):  # the shard holds a window only: halo exchange
                self.exchange = HaloExchange(dist, rank, world, self.n, needs, base=wlo)
                self.exchange_kind = "nccl-halo"

    def step(self):
        if self.world == 1 or self.exchange_kind == "p2p":
            self.solver.solve_async()  # unsharded: one CUDA graph per solve after the first
            return
        backward_induction_sharded(
            self.T, self.n, self.rank, self.world, self.solver.solve_period_async, self.V,
            self.dist.all_gather_into_tensor, self.exchange)

    def close(self):
        self.solver.close()
