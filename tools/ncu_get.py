#!/usr/bin/env python
"""Prints selected raw metrics of an .ncu-rep: ncu_get.py <rep> <regex> [<regex> ...]"""
import csv, re, subprocess, sys
rep, pats = sys.argv[1], [re.compile(p) for p in sys.argv[2:]]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:90] if "Kernel Name" in h else "")
    for k, un, x in zip(h, u, r):
        if x and any(p.search(k) for p in pats):
            print("  %-90s %s %s" % (k, x, un))
