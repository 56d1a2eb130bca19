"""fp64 instruction audit: what libsdpb200 CLAIMS a kernel executes (sdpb_stats.fp64_ops, the numerator of every
roofline fraction bench.py prints) against what ncu COUNTS (DADD + DMUL + DFMA thread instructions).

    python tools/audit_fp64.py list                 # case names
    python tools/audit_fp64.py run <case>           # one solve; prints  CASE <name> <kernel> <evals> <fp64_ops>
    python tools/audit_fp64.py table <dir>          # reads <dir>/audit_<case>.{csv,txt} written by the loop below

on the GPU box, per case:
    python tools/audit_fp64.py run $c > gpurun_out/audit_$c.txt &&
    ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dsetp_pred_on.sum --clock-control none --csv \
        --log-file gpurun_out/audit_$c.csv python tools/audit_fp64.py run $c
"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KERNEL_NAMES = {1: "bi_generic", 2: "bi_inv_tiled", 3: "bi_backorder_staged", 4: "bi_cash_int", 5: "bi_inv_tiled2",
                6: "bi_lead_slab", 7: "bi_lead_col", 8: "bi_cash_diag (+ bi_cash_int in period T)", 9: "bi_lead_q2", 14: "bi_lead_q2m",
                10: "bi_two_product_row", 11: "bi_inv_fused", 12: "bi_cash_row", 13: "bi_cash_tail"}


def cases():
    import sdpb200 as S
    c = S.configs
    pm = S.poisson_pmf
    out = {
        "tiled2_c5": (lambda: c.c5(n_states=200_000, T=3), dict(kernel=S.KERNEL_TILED2)),
        "tiled_c5": (lambda: c.c5(n_states=200_000, T=3), dict(kernel=S.KERNEL_TILED)),
        "fused_c1": (lambda: c.c1(), dict(kernel=S.KERNEL_FUSED)),
        "q2_c4": (lambda: c.c4(T=3, inv_half=40), dict()),
        "q2m_c4": (lambda: c.c4(T=2), dict()),   # the full grid: an unsliced launch, so the shared-product kernel
        "col_c4": (lambda: c.c4(T=3, inv_half=40), dict(kernel=S.KERNEL_LEAD_COL)),
        "slab_c4": (lambda: c.c4(T=3, inv_half=40), dict(kernel=S.KERNEL_LEAD_SLAB)),
        "staged_c4": (lambda: c.c4(T=3, inv_half=40), dict(kernel=S.KERNEL_STAGED)),
        "generic_c4": (lambda: c.c4(T=3, inv_half=10), dict(kernel=S.KERNEL_GENERIC)),
        "generic_a": (lambda: c.c5(n_states=20_000, T=3), dict(kernel=S.KERNEL_GENERIC)),
        "diag_c3": (lambda: c.c3(T=3, inv_max=120, cash_max=1000), dict()),
        "cashint_c3": (lambda: c.c3(T=3, inv_max=120, cash_max=1000), dict(kernel=S.KERNEL_CASH_INT)),
        "generic_c3": (lambda: c.c3(T=3, inv_max=40, cash_max=600), dict(kernel=S.KERNEL_GENERIC)),
        "cashrow_frac": (lambda: S.cash_constraint_model(pm([60.0] * 3), price=10, vari_cost=1, salvage=0.5, max_order=100,
                                                          inv_min=0, inv_max=100, cash_min=0, cash_max=400), dict()),
        "generic_frac": (lambda: S.cash_constraint_model(pm([60.0] * 3), price=10, vari_cost=1, salvage=0.5, max_order=100,
                                                          inv_min=0, inv_max=100, cash_min=0, cash_max=400),
                         dict(kernel=S.KERNEL_GENERIC)),
        "overdraft": (lambda: S.cash_overdraft_model(pm([60.0] * 3), price=10, vari_cost=1, overhead_t=[50] * 3, od_limit=300,
                                                     max_order=100, inv_min=0, inv_max=100, cash_min=-300, cash_max=500), dict()),
        "overdraft_generic": (lambda: S.cash_overdraft_model(pm([60.0] * 3), price=10, vari_cost=1, overhead_t=[50] * 3,
                                                             od_limit=300, max_order=100, inv_min=0, inv_max=100,
                                                             cash_min=-300, cash_max=500), dict(kernel=S.KERNEL_GENERIC)),
        "two_product_row": (lambda: S.two_product_cash_model(
            S.GetPmfMulti([[S.PoissonDist(5)] * 3, [S.PoissonDist(6)] * 3], 0.999, 1).tables(), price=(4.0, 5.0),
            vari_cost=(2.0, 3.0), salvage=(1.0, 1.0), q_bound=12, inv_max=24.0, cash_min=0.0, cash_max=250.0), dict()),
        "staff": (lambda: S.workforce_model([0.5, 0.5, 0.5], max_hire=200, max_x=300), dict()),
    }
    return out


def run(name):
    import sdpb200 as S
    mk, kw = cases()[name]
    spec = mk()
    with S.Solver(spec, **kw) as s:
        s.solve()
        st = s.stats()
    print("CASE", name, st["kernel_used"], repr(st["evals"]), repr(st["fp64_ops"]), st["launches"])


def table(d):
    rows = []
    for name in cases():
        txt, cs = os.path.join(d, f"audit_{name}.txt"), os.path.join(d, f"audit_{name}.csv")
        if not (os.path.exists(txt) and os.path.exists(cs)):
            continue
        line = [ln for ln in open(txt) if ln.startswith("CASE")][0].split()
        kern, evals, claimed = int(line[2]), float(line[3]), float(line[4])
        tot = {"dadd": 0.0, "dmul": 0.0, "dfma": 0.0, "dsetp": 0.0}
        kernels = {}
        lines = [ln for ln in open(cs) if not ln.startswith("==")]
        for r in csv.DictReader(lines):
            m = r.get("Metric Name", "")
            for k in tot:
                if f"op_{k}_pred_on" in m:
                    try:
                        v = float(r["Metric Value"].replace(",", ""))
                    except ValueError:
                        continue
                    tot[k] += v
                    if v:
                        kn = r["Kernel Name"].split("(")[0].split("<")[0].replace("void sdpb::", "")
                        kernels[kn] = kernels.get(kn, 0.0) + (v if k in ("dadd", "dmul") else 0.0)
        measured = tot["dadd"] + tot["dmul"]
        rows.append((name, KERNEL_NAMES.get(kern, str(kern)), evals, claimed / evals, measured / evals, tot["dfma"] / evals,
                     tot["dsetp"] / evals, measured / claimed if claimed else float("nan"),
                     ", ".join(f"{k} {v / measured:.0%}" for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]) if v)))
    print("| case | kernel (as reported) | evaluations | claimed DADD+DMUL / eval | ncu DADD+DMUL / eval | ncu DFMA / eval | "
          "ncu DSETP / eval | measured / claimed | kernels that executed them |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r[0]} | {r[1]} | {r[2]:.4g} | {r[3]:.3f} | {r[4]:.3f} | {r[5]:.4f} | {r[6]:.3f} | {r[7]:.3f} | {r[8]} |")


if __name__ == "__main__":
    if sys.argv[1] == "list":
        print(" ".join(cases()))
    elif sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        table(sys.argv[2])
