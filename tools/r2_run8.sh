#!/bin/bash
# round 2, GPU call 8 (1 GPU): full default bench line (extras: cfg4, cfg5, parity), reference arm, launch list
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 > gpurun_out/r2_bench_h2_full.json 2> gpurun_out/r2_bench_h2_full.err; echo "bench exit $?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_h2_full.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['value'], d['e2e'], d.get('parity'), d.get('cfg4'), d.get('cfg5'), d.get('cpu_baseline'), d.get('e2e_u8'))
    print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('mma_issue_floor'), d['roofline'].get('traffic'))
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/r2_bench_h2_full.err').read()[-3000:])
P
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cut -c1-600 gpurun_out/r2_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.txt 2>&1; head -50 gpurun_out/r2_launches_summary.txt
