"""Generates tests/golden/fullsize.json: the CPU oracle's WHOLE-GRID, FULL-HORIZON result on the
measured configurations C3 (T=12), C4 (T=20) and the C5 bench grid (1e7 states, T=4), frozen as
one SHA-256 per period over the bytes of the float64 value table and of the float64 order-quantity
table, plus V_1 / Q_1 at the initial state.  The GPU tests hash what libsdpb200 produced and
compare -- a whole-grid, every-period, bit-for-bit check that costs one solve and one D2H copy.

The oracle takes minutes per configuration on 8 cores (C3 4.6e11, C4 5.2e11, C5 1.6e12
evaluations), which is why the result is frozen instead of being recomputed on the GPU box.  The pmf
tables the run used are stored next to the hashes so the fixture does not depend on the scipy build.

    python tests/golden/make_fullsize_golden.py [c3] [c4] [c5_1e7]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_lib as O  # noqa: E402
import sdpb200 as S  # noqa: E402

OUT = os.path.join(HERE, "fullsize.json")
CONFIGS = {
    "c3": (lambda: S.configs.c3(), [[0.0, 100.0]]),
    "c4": (lambda: S.configs.c4(), [[0.0, 0.0, 0.0]]),
    "c5_1e7": (lambda: S.configs.c5(n_states=10_000_000), [[0.0]]),
}


def table_hashes(V, Q):
    """One digest per period over the raw little-endian float64 bytes."""
    hv = [hashlib.sha256(np.ascontiguousarray(V[t], dtype="<f8").tobytes()).hexdigest() for t in range(len(V))]
    hq = [hashlib.sha256(np.ascontiguousarray(Q[t], dtype="<f8").tobytes()).hexdigest() for t in range(len(Q))]
    return hv, hq


def main():
    names = sys.argv[1:] or list(CONFIGS)
    done = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        make, init = CONFIGS[name]
        spec = make()
        t0 = time.time()
        V, Q, evals, off = O.dense(spec, threads=0)
        dt = time.time() - t0
        hv, hq = table_hashes(V, Q)
        i0 = O.index(spec, init[0])
        done[name] = {
            "n_states": int(V.shape[1]), "T": int(spec.T), "evals": evals, "offgrid": int(off),
            "init": init[0], "V1_init": float(V[0, i0]).hex(), "V1_init_repr": repr(float(V[0, i0])),
            "Q1_init": float(Q[0, i0]),
            "sha256_V": hv, "sha256_Q": hq,
            "pmf": [np.asarray(r, dtype=np.float64).tolist() for r in spec.pmf[:1]],
            "pmf_same_every_period": all(np.array_equal(np.asarray(r), np.asarray(spec.pmf[0])) for r in spec.pmf),
            "oracle_seconds": round(dt, 1), "oracle_threads": os.cpu_count(),
        }
        json.dump(done, open(OUT, "w"), indent=1)
        print(name, V.shape, f"{evals:.4g} evals in {dt:.0f} s", "V1(init) =", repr(float(V[0, i0])),
              "Q1 =", Q[0, i0], flush=True)


if __name__ == "__main__":
    main()
