#!/usr/bin/env python
"""Join an ncu source page (SASS, execution counts) with nvdisasm -g line info:
static hot-code size and dynamic instruction share per source line bucket.
usage: sass_hot_lines.py <ncu source csv> <nvdisasm -g -c out> <kernel substring> [bucket]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ex = []
for r in rows[2:]:
    try: ex.append((int(r[0], 16), int(r[5]), int(r[40]), int(r[4])))
    except Exception: pass
base = ex[0][0]
exd = {a - base: (e, ni, s) for a, e, ni, s in ex}
key = sys.argv[3]; bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 10
insec = False; cur = ("?", 0); sub = "main"
loc = {}
for l in open(sys.argv[2]):
    if l.startswith("//-----"):
        insec = (".text." in l) and key in l; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'^(\$[^:]+):', l)
    if m: sub = m.group(1).split("$")[-1]; sub = re.sub(r'_ZN\d+_INTERNAL_[0-9a-f]+_\d+_ttmpc_solve_cu_[0-9a-f]+', '', sub)[:28]; continue
    m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/', l)
    if m: loc[int(m.group(1), 16)] = (sub, cur[0], cur[1])
ref = sorted((e for e, _, _ in exd.values()), reverse=True)[1500]
tot = sum(e for e, _, _ in exd.values())
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for off, (e, ni, s) in exd.items():
    sb, f, ln = loc.get(off, ("?", "?", 0))
    k = (sb, f, ln // bucket * bucket)
    a = agg[k]
    a[0] += 1
    if e >= 0.05 * ref: a[1] += 1
    a[2] += e; a[3] += ni
print(f"{'function':28s} {'file':22s} {'lines':>11s} {'static':>7s} {'hot':>6s} {'dyn%':>6s} {'noinst':>7s}")
subtot = collections.defaultdict(lambda: [0, 0, 0, 0])
for k in sorted(agg):
    a = agg[k]
    for i in range(4): subtot[k[0]][i] += a[i]
    if a[1] >= 8 or a[2] / tot > 0.003:
        print(f"{k[0]:28s} {k[1]:22s} {k[2]:5d}-{k[2]+bucket-1:5d} {a[0]:7d} {a[1]:6d} {a[2]/tot*100:6.2f} {a[3]:7d}")
print()
for k, a in subtot.items(): print(f"{k:28s} static {a[0]:6d} hot {a[1]:6d} ({a[1]*16/1024:5.1f} KB) dyn {a[2]/tot*100:6.2f}% noinst {a[3]}")
