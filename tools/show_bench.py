"""Print the interesting part of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
for k in ("n_gpus", "value", "ms_per_step", "kernel", "gpu_launches", "verified_vs_oracle_sample", "verified_vs_golden_hash",
          "verified_vs_unsharded", "device_bytes_per_gpu", "peer_bytes_per_period"):
    print(k, d.get(k))
print("e2e", d["e2e"]["value"], d["e2e"]["phases_s_per_step"])
print("roofline frac", d["roofline"]["frac"], "achieved", d["roofline"]["achieved"], "peak", d["roofline"]["peak"])
if d.get("period_profile"):
    pp = d["period_profile"]
    print("period profile sums (max over ranks): kernels/push/wait ms", pp["sum_max_over_ranks_ms"])
    print("  rank0 per period:", pp["rank0"][:4], "...")
for k, c in d.get("configs", {}).items():
    print(f"{k:10s} {c['solve_ms']:9.3f} ms  {str(c['kernel']):14s} frac {c.get('fp64_frac') or 0:.3f}  n_gpus {c['n_gpus']} "
          f"exch {c.get('exchange')} verified {c.get('verified_vs_unsharded')} bytes/gpu {c.get('device_bytes_per_gpu')}")
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("java_probe"))
