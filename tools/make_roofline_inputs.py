#!/usr/bin/env python
"""profiles/r2_roofline_inputs.json: per workload, what bench.py's `roofline` takes from ncu.

usage: make_roofline_inputs.py <workload> <report.ncu-rep> <evals_per_launch> <summary file under profiles/>

From ONE `ncu --set full` capture of the solve kernel on the workload's seed-1000 batch:
  dram_bytes_per_launch   = dram__bytes_read.sum + dram__bytes_write.sum
  executed_flops_per_eval = (2 * DFMA + DADD + DMUL thread instructions, predicated on) / evaluations
                            of that launch (tools/profile_run.py prints them)
bench.py multiplies the per-evaluation figure with the evaluations of its own timed region."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl, rep, evals, src = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
def num(k):
    u, v = d[k]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    return x * scale
dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
# --set full carries the per-cycle rates (sum over the SMSPs / elapsed cycles): x elapsed cycles = totals
cyc = num("smsp__cycles_elapsed.avg") if "smsp__cycles_elapsed.avg" in d else num("sm__cycles_elapsed.avg")
rate = lambda op: num(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")
fl = (2 * rate("dfma") + rate("dadd") + rate("dmul")) * cyc
path = os.path.join(ROOT, "profiles", "r2_roofline_inputs.json")
cur = json.load(open(path)) if os.path.exists(path) else {}
cur[wl] = {"dram_bytes_per_launch": dram, "executed_flops_per_launch": fl, "evals_per_launch": evals,
           "executed_flops_per_eval": fl / evals, "kernel_ms_alone": num("gpu__time_duration.sum") if d["gpu__time_duration.sum"][0] == "ms" else None,
           "source": src}
json.dump(cur, open(path, "w"), indent=1)
print(json.dumps(cur[wl]))
