#!/usr/bin/env python
"""Lane utilisation per source line from an ncu source page + nvdisasm -g listing:
where does the kernel issue instructions with few active lanes (divergence, lane-0 work)?
usage: sass_waste.py <ncu source csv> <nvdisasm -g -c out> <kernel substring> <iterations>"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ie, iat = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed')
data = []
for r in rows[2:]:
    try: data.append((int(r[ia], 16), int(r[ie]), float(r[iat]), r[1].strip()))
    except Exception: pass
base = data[0][0]
insec = False; cur = ("?", 0); sub = "main"; loc = {}
for l in open(sys.argv[2]):
    if l.startswith("//-----"):
        insec = (".text." in l) and sys.argv[3] in l; continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'^(\$[^:]+):', l)
    if m: sub = m.group(1)[-18:]; continue
    m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/', l)
    if m: loc[int(m.group(1), 16)] = (sub, cur[0], cur[1])
IT = float(sys.argv[4])
tot = sum(e for _, e, _, _ in data)
agg = collections.defaultdict(lambda: [0, 0.0])
for a, e, at, t in data:
    k = loc.get(a - base, ('?', '?', 0))
    agg[k][0] += e; agg[k][1] += e * at
print("instructions per iteration %.0f, average active lanes %.1f" % (tot / IT, sum(v[1] for v in agg.values()) / tot))
print("source lines with < 20 active lanes on average, by instructions per iteration:")
for k, v in sorted(agg.items(), key=lambda z: -z[1][0]):
    if v[0] / IT >= 8 and v[1] / v[0] < 20:
        print(f"  {k[0]:18s} {k[1][:18]:18s} {k[2]:5d}  {v[0]/IT:7.1f} instr/iter  {v[1]/v[0]:5.1f} lanes")
