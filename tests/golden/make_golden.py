"""Generates tests/golden/*.npz: the CPU oracle's outputs on the small cases of tests/cases.py.

The reference (Java + SSJ) cannot run in the build container and ships no golden vectors for
this path, so these fixtures are produced by oracle/sdp_oracle.cpp — the
line-by-line restatement of the Java recursion — and serve as frozen regression pins: the oracle
must keep reproducing them (tests/test_oracle.py) and the CUDA path must match them bit for bit
(tests/test_parity_gpu.py).  Each file stores the pmf table too, so the fixtures do not depend on
the scipy version that built it.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
import oracle_lib as O  # noqa: E402

GOLDEN = ["case_A_small", "case_A_gy", "case_A_twopoint", "case_B1_ref", "case_B2_small", "case_C_rich",
          "case_D_rich", "case_E_small", "case_F_small", "case_XR_small"]


def main():
    for name in GOLDEN:
        spec, init = getattr(cases, name)()
        V, Q, evals, off = O.dense(spec, threads=1)
        rows, iv, _ = O.topdown(spec, init)
        keep = {"V": V, "Q": Q.astype(np.float32), "evals": evals, "topdown_rows": rows, "init": np.asarray(init),
                "init_values": iv, "pmf_len": np.array([len(r) for r in spec.pmf]),
                "pmf": np.concatenate([np.asarray(r) for r in spec.pmf])}
        # E_small and C-like grids are large: keep period 1 and T only for those
        if V.size > 40000:
            keep["V"] = V[[0, -1]]
            keep["Q"] = Q[[0, -1]].astype(np.float32)
            keep["periods"] = np.array([1, spec.T])
        else:
            keep["periods"] = np.arange(1, spec.T + 1)
        np.savez_compressed(os.path.join(HERE, name[5:] + ".npz"), **keep)
        print(name, V.shape, f"{evals:.3g} evals", "V1(init) =", iv)


if __name__ == "__main__":
    main()
