#!/usr/bin/env python3
"""Attributes an ncu source-page dump (--page source --print-source cuda,sass --csv) to device
subroutines: SASS rows are sorted by address, split at RET instructions, and each segment is
labelled by the source function that contributes most of its instructions."""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = []
cur_file, cur_line = None, None
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
        sass.append((int(r[2], 16), cur_file, cur_line, r[3], int(r[iex]), int(r[ismp])))
    except ValueError:
        pass
sass.sort()
# function start lines per file
starts = {}
for f in set(s[1] for s in sass if s[1]):
    try:
        lines = open(f).read().split("\n")
    except OSError:
        continue
    st = []
    for i, l in enumerate(lines, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|__host__ __device__)[^(]*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m and not l.startswith(" "):
            st.append((i, m.group(1)))
    starts[f] = st


def func_of(f, line):
    name = "?"
    for i, n in starts.get(f, []):
        if i <= line:
            name = n
        else:
            break
    return name


tot_ex = sum(s[4] for s in sass)
tot_sm = sum(s[5] for s in sass)
print("total warp-instructions %d, samples %d, sass rows %d" % (tot_ex, tot_sm, len(sass)))
seg, segs = [], []
for s in sass:
    seg.append(s)
    if re.match(r"^\s*(@!?U?P\d+\s+)?RET", s[3]):
        segs.append(seg)
        seg = []
if seg:
    segs.append(seg)
print("%-8s %6s %7s %7s  %s" % ("segment", "instrs", "exec%", "smp%", "dominant source functions (by executed instructions)"))
for k, sg in enumerate(segs):
    ex = sum(s[4] for s in sg)
    sm = sum(s[5] for s in sg)
    if ex == 0:
        continue
    by = collections.Counter()
    for s in sg:
        by[func_of(s[1], s[2])] += s[4]
    label = ", ".join("%s %.0f%%" % (n, 100.0 * c / ex) for n, c in by.most_common(5))
    print("%-8d %6d %6.1f%% %6.1f%%  %s" % (k, len(sg), 100.0 * ex / tot_ex, 100.0 * sm / max(tot_sm, 1), label))
