#!/bin/bash
# round 2, GPU call 13 (1 GPU): resident-weight / N=64-stacked generic kernels (parity, A/B timing), pipelined e2e (test + bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_h2.py tests/test_pipeline_gpu.py -q -x > gpurun_out/r2_pytest_resw.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_resw.log
tail -8 gpurun_out/r2_pytest_resw.log
echo "--- streamed weights (round-2a kernels)"; WCTB_H2_RESIDENT=0 timeout 300 python tools/profile_h2_generic.py 64:64:540:960:0 64:64:540:960:1 32:32:1080:1920:1 32:32:1080:1920:0 16:32:1080:1920:0 32:16:1080:1920:0 32:64:540:960:0 64:32:540:960:2 2>&1 | tee gpurun_out/r2_h2_generic_timing_streamed.txt
echo "--- resident weights"; timeout 300 python tools/profile_h2_generic.py 64:64:540:960:0 64:64:540:960:1 32:32:1080:1920:1 32:32:1080:1920:0 16:32:1080:1920:0 32:16:1080:1920:0 32:64:540:960:0 64:32:540:960:2 2>&1 | tee gpurun_out/r2_h2_generic_timing_resident.txt
if grep -q "pytest exit 0" gpurun_out/r2_pytest_resw.log; then
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_h2_resw.json 2> gpurun_out/r2_bench_h2_resw.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_h2_resw.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], d.get('parity'))
for x in d['roofline']['by_shape']: print('  ',x)
P
tail -5 gpurun_out/r2_bench_h2_resw.err
fi
