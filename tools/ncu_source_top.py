#!/usr/bin/env python3
"""Top source lines by warp-stall samples for one kernel launch of an ncu report (needs -lineinfo and --import-source on).
usage: ncu_source_top.py <report.ncu-rep> <kernel regex> [launch-skip] [top N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, fname, items, tot, tot_inst = None, "", [], 0, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        print("kernel:", r[1]); continue
    if r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr is None or r[0] == "":
        continue  # SASS rows
    try:
        n, ie = int(r[si]), int(r[ii])
    except (ValueError, IndexError):
        continue
    tot += n; tot_inst += ie
    items.append((n, ie, fname, r[0], r[1].strip()[:120]))
items.sort(reverse=True)
print(f"total samples {tot}, warp instructions {tot_inst}")
for n, ie, f, ln, src in items[:top]:
    print(f"{100 * n / max(tot, 1):5.1f}% samples {100 * ie / max(tot_inst, 1):5.1f}% inst  {f}:{ln}  {src}")
