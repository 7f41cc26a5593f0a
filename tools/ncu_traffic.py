#!/usr/bin/env python
"""Read this round's `ncu --set full` captures (gpurun_out/*.ncu-rep) here on the CPU box and write
  profiles/r02_traffic.json   -- DRAM read+write bytes per launch, keyed "<kernel>@<W>x<H>", consumed by bench.py (roofline.traffic)
  profiles/r02_ncu_summary.txt -- the metrics B200_PROFILING.md asks for, one block per capture
usage: python tools/ncu_traffic.py name=rep_path:kernel_key:WxH:algorithmic_bytes ..."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        res.append({h: (v, u) for h, u, v in zip(hdr, units, vals)})
    return res


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def to_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return num(v) * f


traffic_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
lines = []
for spec in sys.argv[1:]:
    name, rest = spec.split("=", 1)
    rep, key, shape, alg = rest.split(":")
    for k in raw(rep):
        kn = k.get("Kernel Name", ("?", ""))[0]
        rd, wr = to_bytes(*k["dram__bytes_read.sum"]), to_bytes(*k["dram__bytes_write.sum"])
        traffic["%s@%s" % (key, shape)] = {"dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr), "algorithmic_bytes": int(float(alg)),
                                           "shape": shape, "source": "ncu --set full, %s (%s)" % (os.path.basename(rep), name)}
        lines.append("== %s: %s  [%s]" % (name, kn[:90], os.path.basename(rep)))
        for w in WANT:
            if w in k:
                lines.append("   %-72s %s %s" % (w, k[w][0], k[w][1]))
        lines.append("   traffic/algorithmic = %.2f" % ((rd + wr) / float(alg)))
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
with open(os.path.join(ROOT, "profiles", "r02_ncu_summary.txt"), "a") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
