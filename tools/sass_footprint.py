#!/usr/bin/env python
"""Attribute the SASS of one kernel to source lines (needs -lineinfo).
usage: sass_footprint.py <nvdisasm -g -c output> <kernel substring> [bucket ranges file:lo-hi=name ...]"""
import re, sys, collections
path, key = sys.argv[1], sys.argv[2]
insec = False; cur = ("?", 0); hist = collections.Counter(); total = 0
sub = "main"; subhist = collections.Counter()
for l in open(path):
    if l.startswith("//-----"):
        insec = (".text." in l) and key in l
        continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'^(\$[^:]+):', l)
    if m:
        n = m.group(1); sub = n.split("$")[-1][:60]
        continue
    if re.match(r'^\s+/\*[0-9a-f]{4,}\*/', l):
        subhist[sub] += 1; total += 1
        if len(sys.argv) < 4 or sys.argv[3] in sub: hist[cur] += 1
print("total instructions", total, "=", total * 16 // 1024, "KB")
for k, v in subhist.most_common(): print(f"  {v:6d} {v*16/1024:6.1f} KB  {k}")
byfile = collections.defaultdict(list)
for (f, ln), v in hist.items(): byfile[f].append((ln, v))
for f, lst in byfile.items():
    lst.sort()
    print(f, sum(v for _, v in lst))
    # 20-line buckets
    b = collections.Counter()
    for ln, v in lst: b[ln // 20 * 20] += v
    for ln in sorted(b): print(f"   {ln:5d}-{ln+19:5d}: {b[ln]:6d}")
