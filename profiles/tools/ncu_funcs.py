#!/usr/bin/env python
"""Warp instructions and stall samples of one kernel from an .ncu-rep, summed per enclosing source function (needs -lineinfo).
usage: python profiles/tools/ncu_funcs.py REPORT KERNEL [top]"""
import csv
import os
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "ima2p_b200", "csrc")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = {}


def func_of(fname, line):
    if fname not in starts:
        s = []
        path = os.path.join(ROOT, fname)
        if os.path.exists(path):
            for n, t in enumerate(open(path), 1):
                m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:IMA_DEV|IMA_HD|IMA_KERNEL|static|inline|__device__)\b[^;{]*?([A-Za-z_][A-Za-z_0-9]*)\s*\(", t)
                if m:
                    s.append((n, m.group(1)))
        starts[fname] = s
    name = "?"
    for n, f in starts[fname]:
        if n <= line:
            name = f
        else:
            break
    return name if starts[fname] else fname


fname, hdr, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            k = (fname, func_of(fname, int(r[0])))
            a = agg.setdefault(k, [0, 0])
            a[0] += int(r[ie]); a[1] += int(r[ss])
        except ValueError:
            pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print("kernel %s: %d warp instructions, %d samples" % (kern, tot_i, tot_s))
for (f, fn), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% smpl  %s: %s" % (100.0 * i / tot_i, 100.0 * s / max(tot_s, 1), f, fn))
