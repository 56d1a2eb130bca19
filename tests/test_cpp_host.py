"""The C++ host-side mirror (include/sdpb200.hpp) through a driver written like the reference's main()s
(tests/cpp_driver.cpp).  CPU: it compiles with -Wall -Werror, links, and refuses to run without a device.
GPU: every number it prints equals the oracle's on the pmf tables the driver itself built."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stochastic-inventory_b200")


def _build(tmp_path):
    exe = tmp_path / "cpp_driver"
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp_driver.cpp"), "-o", str(exe), "-L", PKG, "-lsdpb200",
                    "-Wl,-rpath," + PKG], check=True)
    return exe


def _read_pmfs(path):
    out, lines = {}, open(path).read().split("\n")
    i = 0
    while i < len(lines) and lines[i].strip():
        name, T = lines[i].split()
        rows = []
        for t in range(int(T)):
            f = lines[i + 1 + t].split()
            rows.append(np.array(f[1:], dtype=float).reshape(int(f[0]), 2))
        out[name] = rows
        i += 1 + int(T)
    return out


def test_cpp_host_compiles_and_has_no_cpu_path(tmp_path, S):
    S.abi.load()
    exe = _build(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([str(exe), str(tmp_path / "x.pmf")], capture_output=True, text=True)
        assert r.returncode == 77 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_host_matches_oracle(tmp_path, S, oracle):
    exe = _build(tmp_path)
    pmf_file = tmp_path / "driver.pmf"
    out = subprocess.run([str(exe), str(pmf_file)], capture_output=True, text=True, check=True).stdout
    res = {ln.split()[0]: [float(x) for x in ln.split()[1:]] for ln in out.strip().split("\n")}
    pmf = _read_pmfs(pmf_file)
    # the C++ GetPmf agrees with the scipy mirror to rounding (same support, probabilities within 1e-12)
    ref = S.GetPmf([S.PoissonDist(m) for m in (9, 23, 53, 29)], 0.9999, 1).getpmf()
    for a, b in zip(pmf["A"], ref):
        assert a.shape == b.shape and np.array_equal(a[:, 0], b[:, 0]) and np.allclose(a[:, 1], b[:, 1], rtol=1e-11, atol=0)

    spec = S.inventory_model(pmf["A"], fixed_cost=500, vari_cost=0, hold_cost=2, penalty_cost=10, max_order=60,
                             inv_min=-120, inv_max=200)
    rows, iv, _ = oracle.topdown(spec, [[0.0]])
    assert res["A"] == [iv[0], rows[0][-2], float(len(rows)), 1.0]          # value, Q*, table rows, getAction threw
    assert res["G"] == [iv[0], rows[0][-2]]                                 # the same on three shards (sdpb_group_*)
    spec_b = S.inventory_model(pmf["A"], fixed_cost=500, vari_cost=0, hold_cost=2, penalty_cost=10, max_order=60,
                               inv_min=-120, inv_max=200)
    spec_b.terminal_value = spec_b.tabulate(lambda x: -0.75 * np.maximum(x, 0) + 2.5 * np.maximum(-x, 0))
    rows_b, iv_b, _ = oracle.topdown(spec_b, [[0.0]])
    assert res["H"] == [iv_b[0], rows_b[0][-2]] and iv_b[0] != iv[0]        # with a boundary function
    sweep = []
    for K in (300, 500, 700):                                               # Engine::solveBatch: three engines, one graph
        sp = S.inventory_model(pmf["A"], fixed_cost=K, vari_cost=0, hold_cost=2, penalty_cost=10, max_order=60,
                               inv_min=-120, inv_max=200)
        sweep.append(oracle.topdown(sp, [[0.0]])[1][0])
    assert res["J"] == sweep and sweep[1] == iv[0]
    assert abs(res["K"][0] - iv[0]) <= 1e-9 * abs(iv[0]) and res["K"][1] == rows[0][-2]   # the opt-in collapsed kernel
    assert res["A_row0"] == [rows[0][0], rows[0][1], rows[0][-2]]
    assert res["A_rowN"] == [rows[-1][0], rows[-1][1], rows[-1][-2]]

    spec = S.leadtime_model(pmf["B"], fixed_cost=0, vari_cost=1, hold_cost=2, penalty_cost=10, max_order=12,
                            inv_min=-45, inv_max=36, lead_time=1, clamp=False)
    rows, iv, _ = oracle.topdown(spec, [[0.0, 0.0]])
    assert res["B"] == [iv[0], rows[0][-2]]
    assert res["I"][:2] == [iv[0], rows[0][-2]]                             # grid sized by sdpb_reachable_hull
    assert res["I"][2:] == list(S.reachable_hull(spec, [[0.0, 0.0]]))

    spec = S.cash_constraint_model(pmf["C"], price=8, vari_cost=1, fixed_cost=10, hold_cost=0.5, salvage=0.5, overhead=4,
                                   overhead_rate=0.02, deposit_rate=0.05, penalty_cost=0.3, max_order=15, inv_min=0,
                                   inv_max=30, cash_min=0, cash_max=150, gamma=0.95)
    rows, iv, _ = oracle.topdown(spec, [[0.0, 30.0]])
    assert res["C"] == [iv[0], rows[0][-2], float(len(rows))]

    spec = S.cash_survival_model(pmf["F"], price_t=[4, 5, 4], vari_cost_t=[1, 2, 1], overhead_t=[12, 10, 14], salvage=0.5,
                                 hold_cost=0, deposit_rate=0, fixed_cost=0, max_order=20, inv_min=0, inv_max=40,
                                 cash_min=-30, cash_max=120)
    rows, iv, _ = oracle.topdown(spec, [[0.0, 15.0]])
    assert res["F"] == [iv[0], rows[0][-2]]
    assert 0.0 < iv[0] <= 1.0
