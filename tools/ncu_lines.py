#!/usr/bin/env python3
"""Rank CUDA source lines of an ncu report by warp-stall samples / executed instructions.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > mix.csv ; tools/ncu_lines.py mix.csv [N]"""
import csv
import sys

def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif len(r) > 2 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 10 and r[0].isdigit():
        d = dict(zip(hdr, r))
        out.append((cur, int(r[0]), num(d["Instructions Executed"]), num(d["# Samples"]), r[1].strip()[:90]))
tot = sum(o[2] for o in out) or 1
ts = sum(o[3] for o in out) or 1
print("total warp instructions", tot, "samples", ts)
for o in sorted(out, key=lambda x: -x[3])[:top]:
    print("%-20s %4d inst %5.1f%% samp %5.1f%%  %s" % (o[0], o[1], 100 * o[2] / tot, 100 * o[3] / ts, o[4]))
