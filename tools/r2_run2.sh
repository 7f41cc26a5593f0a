#!/bin/bash
# round 2, GPU call 2: h2 engine -- rates, full h2 test file, first bench lines (h2 vs tf32 on the same box), timeline
mkdir -p gpurun_out
timeout 300 python tools/h2_rates.py > gpurun_out/r2_h2_rates.txt 2>&1; echo "exit $?" >> gpurun_out/r2_h2_rates.txt
timeout 1200 python -m pytest tests/test_gpu_h2.py -q -s > gpurun_out/r2_pytest_h2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_h2.log
timeout 600 python bench.py --precision h2 --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_h2_unfused.json 2> gpurun_out/r2_bench_h2_unfused.err
timeout 600 python bench.py --precision tf32 --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_tf32.json 2> gpurun_out/r2_bench_tf32.err
timeout 300 python tools/stage_timeline.py h2 > gpurun_out/r2_timeline_h2_unfused.txt 2>&1
timeout 300 python tools/stage_timeline.py tf32 > gpurun_out/r2_timeline_tf32.txt 2>&1
cat gpurun_out/r2_h2_rates.txt; grep -E "cfg|passed|failed|h2 stage" gpurun_out/r2_pytest_h2.log | tail -40
cat gpurun_out/r2_bench_h2_unfused.json | cut -c1-400; tail -3 gpurun_out/r2_bench_h2_unfused.err; cat gpurun_out/r2_bench_tf32.json | cut -c1-300
cat gpurun_out/r2_timeline_h2_unfused.txt gpurun_out/r2_timeline_tf32.txt
