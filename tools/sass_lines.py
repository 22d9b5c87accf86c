#!/usr/bin/env python
"""128-byte instruction-cache lines touched at least `thr` times per PANOC iteration, per function.
usage: sass_lines.py <ncu source csv> <iterations> [thr=0.3]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
IT = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
ex = []
for r in rows[2:]:
    try: ex.append((int(r[0], 16), int(r[5])))
    except Exception: pass
base = ex[0][0]
lines = collections.defaultdict(lambda: [0, 0, 0])
for a, e in ex:
    l = (a - base) // 128
    L = lines[l]; L[0] += 1; L[1] = max(L[1], e); L[2] += (e >= thr * IT)
hot = [l for l, L in lines.items() if L[1] >= thr * IT]
print("lines touched >= %.2f/iter: %d = %.1f KB; hot instructions in them %d (fill %.0f%%)" % (
    thr, len(hot), len(hot) * 128 / 1024, sum(lines[l][2] for l in hot), 100 * sum(lines[l][2] for l in hot) / (8.0 * len(hot))))
dyn = sum(e for _, e in ex) / IT
print("dynamic instructions per iteration %.0f -> %.0f line requests if nothing is reused" % (dyn, dyn / 8))
# contiguous runs of hot lines
runs = []; cur = None
for l in sorted(hot):
    if cur and l == cur[1] + 1: cur[1] = l
    else:
        cur = [l, l]; runs.append(cur)
print("runs of hot lines:", len(runs), " median run %d lines" % sorted(r[1] - r[0] + 1 for r in runs)[len(runs) // 2])
for r in runs:
    if r[1] - r[0] + 1 >= 6: print("   +%5d..+%5d instrs  %5.1f KB" % (r[0] * 8, r[1] * 8 + 7, (r[1] - r[0] + 1) / 8.0))
