"""`Solver`: one sdpb_handle.  Thin, typed access to the C-ABI; all compute happens in
libsdpb200.so on the GPU.  Nothing here can solve anything on the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A
from .models import ModelSpec


class Solver:
    def __init__(self, spec: ModelSpec, device: int = -1, shard_rank: int = 0, shard_count: int = 1,
                 kernel: int = A.KERNEL_AUTO, dedup: bool = False, stream: int | None = None):
        self.lib = A.load()
        self.spec = spec
        self._model = spec.to_struct()
        opt = A.SdpbOptions()
        opt.struct_size = C.sizeof(A.SdpbOptions)
        opt.device, opt.shard_rank, opt.shard_count = device, shard_rank, shard_count
        opt.kernel, opt.dedup = kernel, 1 if dedup else 0
        opt.stream = stream
        h = C.c_void_p()
        rc = self.lib.sdpb_create(C.byref(self._model), C.byref(opt), C.byref(h))
        if rc != A.SDPB_OK:
            raise A.SdpbError(rc, self.lib.sdpb_last_error(None).decode())
        self.h = h
        g = A.SdpbGrid()
        self._check(self.lib.sdpb_grid_info(self.h, C.byref(g)))
        self.grid = g
        self.ndim = g.ndim
        self.n_states = g.n_states
        self.T = g.T

    def _check(self, rc):
        if rc != A.SDPB_OK:
            raise A.SdpbError(rc, self.lib.sdpb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.sdpb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- solve ----
    def solve(self):
        self._check(self.lib.sdpb_solve(self.h))
        return self

    def solve_async(self):
        """Enqueue the whole solve on the handle's stream (CUDA-graph replay from the second call)."""
        self._check(self.lib.sdpb_solve_async(self.h))

    def solve_period_async(self, period: int):
        self._check(self.lib.sdpb_solve_period_async(self.h, period))

    def sync(self):
        self._check(self.lib.sdpb_sync(self.h))

    # ---- results ----
    def value(self, period: int, states):
        """(V, Q) for an array of API-order states [n, ndim]."""
        st = np.ascontiguousarray(np.atleast_2d(np.asarray(states, dtype=np.float64)))
        if st.shape[1] != self.ndim:
            raise ValueError(f"states must have {self.ndim} columns")
        n = st.shape[0]
        v = np.empty(n)
        q = np.empty(n)
        dp = C.POINTER(C.c_double)
        self._check(self.lib.sdpb_value(self.h, period, st.ctypes.data_as(dp), n,
                                        v.ctypes.data_as(dp), q.ctypes.data_as(dp)))
        return v, q

    def period_tables(self, period: int, want_q: bool = True, out_v=None, out_q=None):
        """Whole-grid V_t and order quantities.  `out_v` / `out_q`: caller-owned float64 arrays of n_states
        (e.g. views of page-locked memory, which the device copies into at PCIe speed)."""
        dp = C.POINTER(C.c_double)
        for o in (out_v, out_q):
            if o is not None and (o.dtype != np.float64 or o.size != self.n_states or not o.flags.c_contiguous):
                raise ValueError("output buffers must be contiguous float64 arrays of n_states")
        V = out_v if out_v is not None else np.empty(self.n_states)
        Q = (out_q if out_q is not None else np.empty(self.n_states)) if want_q else None
        self._check(self.lib.sdpb_period_tables(self.h, period, V.ctypes.data_as(dp),
                                                Q.ctypes.data_as(dp) if want_q else None))
        return V, Q

    def device_tables(self, period: int):
        dv, dq = C.c_void_p(), C.c_void_p()
        self._check(self.lib.sdpb_device_tables(self.h, period, C.byref(dv), C.byref(dq)))
        return dv.value, dq.value

    def shard_reads(self):
        """[lo, hi) of the flattened V_{t+1} this shard's kernels may read (multi-GPU halo exchange)."""
        lo, hi = C.c_int64(), C.c_int64()
        self._check(self.lib.sdpb_shard_reads(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def state_of_index(self, idx: int):
        st = np.empty(self.ndim)
        self._check(self.lib.sdpb_state_of_index(self.h, idx, st.ctypes.data_as(C.POINTER(C.c_double))))
        return st

    def reach(self, init_states):
        st = np.ascontiguousarray(np.atleast_2d(np.asarray(init_states, dtype=np.float64)))
        self._check(self.lib.sdpb_reach(self.h, st.ctypes.data_as(C.POINTER(C.c_double)), st.shape[0]))

    def opt_table(self):
        n = C.c_size_t(0)
        self._check(self.lib.sdpb_opt_table(self.h, None, C.byref(n)))
        rows = np.empty((n.value, self.ndim + 2))
        self._check(self.lib.sdpb_opt_table(self.h, rows.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return rows

    def simulate(self, init_state, samples, discount: float = 1.0):
        """Roll the solved policy over demand sample paths [n, T] -> per-path sums (Simulation.java:59-70)."""
        st = np.ascontiguousarray(init_state, dtype=np.float64)
        sm = np.ascontiguousarray(samples, dtype=np.float64)
        if sm.ndim != 2 or sm.shape[1] != self.T:
            raise ValueError(f"samples must be [n, {self.T}]")
        vals = np.empty(sm.shape[0])
        dp = C.POINTER(C.c_double)
        self._check(self.lib.sdpb_simulate(self.h, st.ctypes.data_as(dp), sm.ctypes.data_as(dp), sm.shape[0],
                                           float(discount), vals.ctypes.data_as(dp)))
        return vals

    def stats(self):
        s = A.SdpbStats()
        self._check(self.lib.sdpb_stats_get(self.h, C.byref(s)))
        return {"evals": s.evals, "solve_ms": s.solve_ms, "kernel_ms": s.kernel_ms,
                "launches": s.launches, "kernel_used": s.kernel_used, "fp64_ops": s.fp64_ops,
                "evals_executed": s.evals_executed}
