#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep (needs -lineinfo).
usage: python profiles/tools/ncu_lines.py REPORT KERNEL [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            agg[(fname, int(r[0]))] = (int(r[ie]), int(r[ss]), r[1].strip()[:110])
        except ValueError:
            pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print("kernel %s: %d warp instructions, %d samples" % (kern, tot_i, tot_s))
for (f, ln), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smpl  %s:%d  %s" % (100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), f, ln, src))
