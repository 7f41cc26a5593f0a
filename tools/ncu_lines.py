#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples from an `ncu --import-source on` report.
usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg, smp, txt = collections.Counter(), collections.Counter(), {}
cur, hdr, tot, tots = None, None, 0, 0
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    try:
        ie, sm = int(d["Instructions Executed"]), int(d["# Samples"])
    except Exception:
        continue
    k = (cur, int(r[0]))
    agg[k] += ie; smp[k] += sm; tot += ie; tots += sm; txt[k] = r[1].strip()[:110]
print("total warp instructions %d, samples %d" % (tot, tots))
for k, v in agg.most_common(topn):
    print("%-18s %5d  inst %5.2f%%  stall-samples %5.2f%%  | %s" % (k[0], k[1], 100.0 * v / tot, 100.0 * smp[k] / max(1, tots), txt[k]))
