#!/usr/bin/env python
"""Joins an `ncu --page source --csv` SASS dump with nvdisasm line info to give per-source-line shares of
stall samples and executed instructions.
usage: ncu_by_line.py <sass.csv> <cubin> <kernel-name-substring> [min_pct]"""
import csv, re, subprocess, sys
from collections import defaultdict
sass_csv, cubin, kname = sys.argv[1:4]
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infn, off = [], None, False, 0
chain, fresh = [], True
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        infn = kname in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        # nvdisasm -gi prints the inline chain innermost first, outermost (the line in the kernel body) last
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        if fresh: chain, fresh = [], False
        chain.append(loc)
        cur = (chain[0][0], chain[0][1], chain[-1] if len(chain) > 1 else None)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2)))
        fresh = True
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > iI and r[0].startswith("0x")]
base = int(body[0][0], 16)
bymap = {o: c for o, c, _ in lines}
agg = defaultdict(lambda: [0.0, 0.0])
ts = ti = 0.0
for r in body:
    o = int(r[0], 16) - base
    c = bymap.get(o)
    key = c[:2] if c else ("?", 0)
    top = c[2] if c and c[2] else key          # attribute inlined code to the call site in the kernel file too
    s, n = float(r[iS] or 0), float(r[iI] or 0)
    agg[("line",) + key][0] += s; agg[("line",) + key][1] += n
    if top != key:
        agg[("site",) + top][0] += s; agg[("site",) + top][1] += n
    ts += s; ti += n
print("total samples %.0f, warp instructions %.0f" % (ts, ti))
for k, (s, n) in sorted(agg.items(), key=lambda kv: (kv[0][1], kv[0][2], kv[0][0])):
    if 100 * s / ts >= minp or 100 * n / ti >= minp:
        print("%-5s %-18s:%-4d samples %5.1f%%  inst %5.1f%%" % (k[0], k[1], k[2], 100 * s / ts, 100 * n / ti))
