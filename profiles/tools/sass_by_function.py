#!/usr/bin/env python
"""Static SASS instruction count of one kernel, summed per enclosing source function (from `nvdisasm -g` output of the cubin).
usage: cuobjdump -xelf all lib.so; nvdisasm -g X.cubin > all.sass; python profiles/tools/sass_by_function.py all.sass KERNEL_MANGLED_SUBSTRING"""
import collections
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "ima2p_b200", "csrc")
starts = {}


def func_of(path, line):
    fname = os.path.basename(path)
    if fname not in starts:
        s = []
        p = os.path.join(ROOT, fname)
        if os.path.exists(p):
            for n, t in enumerate(open(p), 1):
                m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:IMA_DEV|IMA_HD|IMA_KERNEL|static|inline|__device__)\b[^;{]*?([A-Za-z_][A-Za-z_0-9]*)\s*\(", t)
                if m:
                    s.append((n, m.group(1)))
        starts[fname] = s
    name = None
    for n, f in starts[fname]:
        if n <= line:
            name = f
        else:
            break
    return "%s:%s" % (fname, name) if name else fname


sass, kern = sys.argv[1], sys.argv[2]
inside, cur, cnt = False, "?", collections.Counter()
for l in open(sass, errors="replace"):
    if l.startswith(".text.") or l.lstrip().startswith(".section"):
        inside = kern in l and ".text." in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = func_of(m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        cnt[cur] += 1
tot = sum(cnt.values())
print("%s: %d SASS instructions (%.0f KB)" % (kern, tot, tot * 16 / 1024))
for k, v in cnt.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 25):
    print("%7d %5.1f%%  %s" % (v, 100.0 * v / max(tot, 1), k))
