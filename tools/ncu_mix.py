"""Instruction mix and headline counters of one kernel from an ncu report: python tools/ncu_mix.py file.ncu-rep [hot]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[-1]
d = dict(zip(hdr, vals))
for k in ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
          "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
          "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"):
    print(f"{k}: {d.get(k)}")
for k in sorted(d):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k] or 0) > 0.05:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]}: {float(d[k]):.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
iS, iE = h.index("Source"), h.index("Instructions Executed")
tot = collections.Counter()
for r in rows[2:]:
    if len(r) <= iE:
        continue
    t = r[iS].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("F2I", "I2F", "F2F", "DSETP", "LDS", "LDG")) else op.split(".")[0]
    tot[op] += int(r[iE] or 0)
total = sum(tot.values())
print("warp instructions:", total)
for op, n in tot.most_common(28):
    print(f"  {op:14s} {100 * n / total:6.2f}%")
