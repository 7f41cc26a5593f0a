#!/bin/bash
# round 2, GPU call 5: fused head/tail with two epilogue groups, packed conversions
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_h2.py -q -x -k "fused or layout" > gpurun_out/r2_pytest_fused.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_fused.log
tail -5 gpurun_out/r2_pytest_fused.log
if grep -q "pytest exit 0" gpurun_out/r2_pytest_fused.log; then
timeout 300 python tools/profile_fused.py > gpurun_out/r2_fused_timing.txt 2>&1; cat gpurun_out/r2_fused_timing.txt
timeout 900 python -m pytest tests/test_gpu_h2.py -q > gpurun_out/r2_pytest_h2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_h2.log
tail -3 gpurun_out/r2_pytest_h2.log
timeout 600 python bench.py --precision h2 --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_h2_fused2.json 2> gpurun_out/r2_bench_h2_fused2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_h2_fused2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'])
for x in d['roofline']['by_shape']: print('  ',x)
P
WCTB_FAST_STATS_H2=1 timeout 600 python bench.py --precision h2 --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fast stats:', d['ms_per_step'])"
timeout 300 python tools/stage_timeline.py h2 > gpurun_out/r2_timeline_h2_fused.txt 2>&1; cat gpurun_out/r2_timeline_h2_fused.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_head_h2 -s 2 -c 1 -o gpurun_out/r2_head_h2_v2 python tools/profile_fused.py --once > gpurun_out/r2_ncu_head.log 2>&1
fi
