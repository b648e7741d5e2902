#!/usr/bin/env python
"""Per-instruction execution counts of one kernel from an ncu report (--import-source on):
prints the SASS with executed warp-instruction counts so hot regions can be read off.

    python tools/ncu_sass_hot.py gpurun_out/prof.ncu-rep [min_count]
"""
import csv
import subprocess
import sys

path = sys.argv[1]
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout.splitlines()
kern = 0
rows = []
for i, line in enumerate(raw):
    if line.startswith('"Kernel Name"'):
        kern += 1
        if kern > 1:
            break
        continue
    rows.append(line)
rd = list(csv.reader(rows))
hdr = rd[0]
ie, src, st = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in rd[1:] if r[ie].isdigit())
print("# total warp instructions executed: %d over %d SASS lines" % (tot, len(rd) - 1))
for k, r in enumerate(rd[1:]):
    print("%5d %9s %6s  %s" % (k, r[ie], r[st], r[src].strip()))
