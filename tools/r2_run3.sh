#!/bin/bash
# round 2, GPU call 3: fused h2 head (dx-stacked, block-pipelined)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_h2.py -q -x -k "fused_head" > gpurun_out/r2_pytest_head.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_head.log
tail -25 gpurun_out/r2_pytest_head.log
if grep -q "pytest exit 0" gpurun_out/r2_pytest_head.log; then
timeout 900 python -m pytest tests/test_gpu_h2.py -q -s > gpurun_out/r2_pytest_h2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_h2.log
grep -E "cfg|passed|failed" gpurun_out/r2_pytest_h2.log | tail
timeout 600 python bench.py --precision h2 --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_h2_head.json 2> gpurun_out/r2_bench_h2_head.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_h2_head.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'])
for x in d['roofline']['by_shape']: print('  ',x)
P
fi
