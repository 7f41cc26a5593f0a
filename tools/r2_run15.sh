#!/bin/bash
# round 2, GPU call 15 (1 GPU): generic h2 kernel with 1 KB inner TMA boxes (parity, timing, bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_h2.py tests/test_gpu_parity.py -q -x > gpurun_out/r2_pytest_tma1k.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_tma1k.log
tail -6 gpurun_out/r2_pytest_tma1k.log
timeout 300 python tools/profile_h2_generic.py 64:64:540:960:0 64:64:540:960:1 32:32:1080:1920:1 32:32:1080:1920:0 16:32:1080:1920:0 32:16:1080:1920:0 32:64:540:960:0 64:32:540:960:2 128:128:270:480:0 64:128:270:480:0 2>&1 | tee gpurun_out/r2_h2_generic_timing_tma1k.txt
if grep -q "pytest exit 0" gpurun_out/r2_pytest_tma1k.log; then
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_h2_tma1k.json 2> gpurun_out/r2_bench_h2_tma1k.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_h2_tma1k.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('parity'), d.get('cfg4'), d.get('cfg5'))
for x in d['roofline']['by_shape']: print('  ',x)
P
tail -5 gpurun_out/r2_bench_h2_tma1k.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_h2_kernel -c 1 -o gpurun_out/r2_h2_64_tma1k python tools/profile_h2_generic.py --once 64:64:540:960:0 > gpurun_out/r2_ncu_h2_64_tma1k.log 2>&1
fi
