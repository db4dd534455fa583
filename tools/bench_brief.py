"""One line per workload of a bench.py JSON line: ms/step, body-steps/s, solver ms and roofline fraction."""
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def row(name, r):
    rf = r["roofline"]
    print(f"{name:16s} {r['ms_per_step']:8.3f} ms/step  {r['value'] / 1e6:9.2f} M body-steps/s  e2e {r['e2e']['value'] / 1e6:9.2f} M  solve {rf['kernel_ms']:.3f} ms  frac {rf['frac']:.3f}"
          + (f"  cpu {r['cpu_baseline']['value'] / 1e3:.1f} k" if "cpu_baseline" in r else ""))
row(j["config"]["workload"] + f" x{j['n_gpus']}", j)
for k, r in j.get("workloads", {}).items():
    row(k, r)
h = j["config"].get("handover")
if h:
    print("handover:", {k: h[k] for k in ("bodies_in_per_rank", "rank0_handover_round_ms", "rank0_handover_phases_ms", "rank0_host_ms_each_step", "rank0_last_step_device_ms") if k in h})
