#!/usr/bin/env python
"""Hot-code footprint of a kernel from an ncu source page (ncu -i rep --page source --csv --print-source sass).
Prints, for execution-count thresholds, how many SASS bytes are executed at least that often, and the
address ranges of the hot code."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ie, isamp, ini = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Warp Stall Sampling (Not-issued Samples)")
inoi = [i for i, h in enumerate(hdr) if "no_instruction" in h.lower() or "No Instruction" in h]
data = []
for r in rows[2:]:
    try: data.append((int(r[ia], 16), int(r[ie]), int(r[isamp]), r[1].strip()))
    except Exception: pass
base = data[0][0]
tot = sum(d[1] for d in data); samp = sum(d[2] for d in data)
print("instructions", len(data), "executed", tot, "samples", samp)
mx = max(d[1] for d in data)
for frac in (0.5, 0.2, 0.1, 0.05, 0.02, 0.01, 0.001):
    th = None
    # threshold on executions relative to the per-iteration count (median of top 2000)
    top = sorted((d[1] for d in data), reverse=True)
    ref = top[1500]
    sel = [d for d in data if d[1] >= ref * frac]
    print(f"exec >= {frac:5.3f} x ref({ref}): {len(sel):6d} instrs = {len(sel)*16/1024:6.1f} KB, {sum(d[1] for d in sel)/tot*100:5.1f}% of executed")
# ranges
ref = sorted((d[1] for d in data), reverse=True)[1500]
blocks = []; cur = None
for a, e, s, t in data:
    hot = e >= 0.05 * ref
    if hot:
        if cur and a - cur[1] <= 16 * 8: cur[1] = a; cur[2] += e; cur[3] += s; cur[4] += 1
        else:
            cur = [a, a, e, s, 1]; blocks.append(cur)
print("hot ranges (offset, KB, instrs hot, exec share, sample share):")
for b in blocks:
    if b[4] >= 16:
        print(f"  +{(b[0]-base)//16:6d} .. +{(b[1]-base)//16:6d}  {(b[1]-b[0]+16)/1024:6.1f} KB  {b[4]:5d}  {b[2]/tot*100:5.1f}%  {b[3]/samp*100:5.1f}%")
