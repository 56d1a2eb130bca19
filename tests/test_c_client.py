"""A plain-C client of the C-ABI: compiles and links on any box (CPU test), runs on the GPU box."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stochastic-inventory_b200")


def _build(tmp_path):
    exe = tmp_path / "c_abi_smoke"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", str(exe), "-L", PKG, "-lsdpb200",
                    "-Wl,-rpath," + PKG], check=True)
    return exe


def test_c_client_links(tmp_path, S):
    S.abi.load()
    exe = _build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode == 77 and "no CUDA device" in r.stderr   # no CPU fallback, from C as well


@pytest.mark.gpu
def test_c_client_matches_oracle(tmp_path, S, oracle):
    exe = _build(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    pmf = [np.array([[1, .2], [2, .5], [3, .3]], dtype=float)] * 2
    spec = S.inventory_model(pmf, fixed_cost=4, vari_cost=1, hold_cost=1, penalty_cost=5, max_order=4,
                             inv_min=-6, inv_max=6)
    rows, iv, _ = oracle.topdown(spec, [[0.0]])
    _, _, evals, _ = oracle.dense(spec)
    assert float(out[0]) == iv[0] and float(out[1]) == rows[0][-2]
    assert int(out[2]) == len(rows)
    assert int(out[3]) == S.abi.SDPB_ERR_UNSOLVED
    assert float(out[4]) == evals
