#!/bin/bash
# round 2, GPU call 19 (1 GPU): fused head / tail with the three hi/lo products in the same 48 TMEM columns (parity, timing, bench)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_h2.py -q -x -k "fused or head or tail or cfg2 or golden" > gpurun_out/r2_pytest_f48.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_f48.log
tail -4 gpurun_out/r2_pytest_f48.log
timeout 300 python tools/profile_fused.py 2>&1 | tee gpurun_out/r2_fused_timing_f48.txt
if grep -q "pytest exit 0" gpurun_out/r2_pytest_f48.log; then
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_h2_f48.json 2> gpurun_out/r2_bench_h2_f48.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_h2_f48.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])
for x in d['roofline']['by_shape'][:4]: print('  ',x)
P
fi
