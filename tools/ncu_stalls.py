#!/usr/bin/env python3
"""Where do the warps of a kernel wait?  SASS-level stall samples of one kernel of an .ncu-rep captured with
--import-source on: the instructions with the most samples and the dominant stall reason at each, plus totals per reason.
    python tools/ncu_stalls.py file.ncu-rep <kernel regex> [top N]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
for b in blocks:
    hdr, rows = b["rows"][0], b["rows"][1:]
    si = hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[si]) for r in rows)
    print("==", b["name"][:110], "samples", total)
    agg = {h: sum(int(r[i]) for r in rows) for i, h in stall_cols}
    print("   per reason:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(1, total)) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ranked = sorted(range(len(rows)), key=lambda k: -int(rows[k][si]))[:top]
    for k in sorted(ranked):
        r = rows[k]
        best = max(stall_cols, key=lambda ih: int(r[ih[0]]))
        print("   %5d  %5.1f%%  %-14s %s" % (k, 100.0 * int(r[si]) / max(1, total), best[1][6:], r[1].strip()[:90]))
