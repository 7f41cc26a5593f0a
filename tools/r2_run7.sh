#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_h2.py -q -x -k "fused or five_stage or cfg2" > gpurun_out/r2_pytest_fused.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_fused.log
tail -4 gpurun_out/r2_pytest_fused.log
timeout 300 python tools/profile_fused.py > gpurun_out/r2_fused_timing.txt 2>&1; cat gpurun_out/r2_fused_timing.txt
for st in 1 0; do
WCTB_STAGGER=$st timeout 600 python bench.py --precision h2 --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_h2_stagger$st.json 2> gpurun_out/r2_bench_h2_stagger$st.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_h2_stagger$st.json').read().strip().splitlines()[-1])
print('stagger=$st', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
P
done
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu_all.log; tail -5 gpurun_out/r2_pytest_gpu_all.log
