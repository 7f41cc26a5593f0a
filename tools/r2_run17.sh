#!/bin/bash
# round 2, GPU call 17 (1 GPU): generic h2 kernel with the warp-cooperative cp.async border loader (parity, timing at three sizes, bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_h2.py tests/test_gpu_parity.py -q -x > gpurun_out/r2_pytest_border.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_border.log
tail -6 gpurun_out/r2_pytest_border.log
timeout 300 python tools/profile_h2_generic.py 64:64:270:480:0 64:64:540:960:0 64:64:1080:1920:0 64:64:540:960:1 32:32:1080:1920:1 32:32:1080:1920:0 16:32:1080:1920:0 32:16:1080:1920:0 32:64:540:960:0 64:32:540:960:2 128:128:270:480:0 128:128:135:240:0 64:128:270:480:0 24:16:2160:3840:0 2>&1 | tee gpurun_out/r2_h2_generic_timing_border.txt
if grep -q "pytest exit 0" gpurun_out/r2_pytest_border.log; then
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_h2_border.json 2> gpurun_out/r2_bench_h2_border.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_h2_border.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('parity'))
print({k: d[k].get('ms_per_step') for k in ('cfg4','cfg5') if k in d})
for x in d['roofline']['by_shape']: print('  ',x)
P
tail -5 gpurun_out/r2_bench_h2_border.err
fi
