#!/usr/bin/env python3
"""Attributes an ncu source-page dump of the cooperative kernel to phases of pmg_coop.cuh.

usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv; ncu_regions.py x.csv
SASS rows are walked in address order; instructions of inlined helpers (dot, cross, shuffles) are charged
to the last pmg_coop.cuh line seen before them, and lines are binned by the `// N.` phase comments."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
sass, cur_file, cur_line = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] in ("Function Name", "Line No") or len(r) <= iex:
        continue
    if r[2] == "-":
        try:
            cur_line = int(r[0])
        except ValueError:
            pass
        continue
    try:
        sass.append((int(r[2], 16), cur_file or "", cur_line, r[3], int(r[iex]), int(r[ismp])))
    except ValueError:
        pass
sass.sort()
coop = [f for f in set(s[1] for s in sass) if f.endswith("pmg_coop.cuh")]
src = open(coop[0]).read().split("\n") if coop else []
# phase markers: function starts and numbered comments
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*>\s*)?__device__[^(]*?\b([A-Za-z_0-9]+)\s*\(", l)
    if m and not l.startswith(" "):
        marks.append((i, m.group(1)))
    m = re.match(r"^\s*// (\d+b?\.|contact rows|projected Gauss|collision detection)(.*)", l)
    if m:
        marks.append((i, "  " + (m.group(1) + m.group(2))[:60]))


def region(line):
    name = "?"
    for i, n in marks:
        if i <= line:
            name = n
        else:
            break
    return name


tot = sum(s[4] for s in sass)
tsm = sum(s[5] for s in sass)
agg, asm, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
last = None
for addr, f, line, text, ex, sm in sass:
    if f.endswith("pmg_coop.cuh") and line:
        last = line
    key = region(last) if last else "(prologue)"
    agg[key] += ex
    asm[key] += sm
    ops[key][text.split()[0].split(".")[0] if not text.startswith("@") else text.split()[1].split(".")[0]] += ex
print("total warp-instructions %d, samples %d" % (tot, tsm))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    top = ", ".join("%s %.0f%%" % (o, 100 * c / v) for o, c in ops[k].most_common(4)) if v else ""
    print("%5.1f%% exec %5.1f%% smp  %-62s %s" % (100 * v / tot, 100 * asm[k] / max(1, tsm), k, top))
