#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: summarize_launches.py launches.csv [last_n_launches]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v * scale))
    if len(sys.argv) > 2:
        rows = rows[-int(sys.argv[2]):]
    agg = defaultdict(lambda: [0, 0.0])
    for name, ms in rows:
        name = re.sub(r"\(.*", "", name)
        agg[name][0] += 1
        agg[name][1] += ms
    total = sum(v[1] for v in agg.values())
    print("launches: %d   total device time: %.3f ms" % (len(rows), total))
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%9.3f ms  %5.1f%%  x%-5d %s" % (ms, 100 * ms / total, n, name[:110]))


if __name__ == "__main__":
    main()
