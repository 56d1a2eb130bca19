"""A plain-C client of the C-ABI: compiles and links on any box (CPU test), runs on the GPU box."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stochastic-inventory_b200")


def _build(tmp_path, name="c_abi_smoke"):
    exe = tmp_path / name
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", name + ".c"), "-o", str(exe), "-L", PKG, "-lsdpb200",
                    "-Wl,-rpath," + PKG], check=True)
    return exe


def test_c_client_links(tmp_path, S):
    S.abi.load()
    exe = _build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode == 77 and "no CUDA device" in r.stderr   # no CPU fallback, from C as well


def test_c_group_client_links(tmp_path, S):
    S.abi.load()
    exe = _build(tmp_path, "c_group_client")
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode == 77 and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("shards", [1, 2, 3, 5])
def test_c_group_client_matches_oracle(shards, tmp_path, S, oracle):
    """A multi-shard solve driven from plain C through sdpb_group_* (no Python, torch or NCCL in the data path)."""
    exe = _build(tmp_path, "c_group_client")
    out = subprocess.run([str(exe)] + ["0"] * shards, capture_output=True, text=True, check=True, timeout=120).stdout.split()
    row = np.array([[j, p] for j, p in enumerate([0.1, 0.2, 0.3, 0.2, 0.15, 0.05])], dtype=float)
    spec = S.leadtime_model([row] * 4, fixed_cost=0, vari_cost=1, hold_cost=2, penalty_cost=10, max_order=7,
                            inv_min=-15, inv_max=15, lead_time=2, clamp=True)
    Vo, Qo, evals, _ = oracle.dense(spec)
    i0 = oracle.index(spec, [0.0, 0.0, 0.0])
    assert float(out[0]) == Vo[0][i0] and float(out[1]) == Qo[0][i0]
    assert float(out[2]) == float(np.cumsum(Vo[0])[-1]) and float(out[3]) == Qo[0].sum()   # same summation order
    assert float(out[4]) == evals and int(out[5]) == Vo.shape[1]
    if shards >= 3:
        assert int(out[6]) < int(out[7])      # a shard holds a window of the tables, not all of them


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
