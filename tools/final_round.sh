#!/bin/bash
# last GPU slot of the round: official-shape bench line (no CPU leg), I/O row timing, then as much of the GPU suite as fits
mkdir -p gpurun_out
timeout 60 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.json | head -c 300; echo
timeout 25 python tools/io_bench.py 2>&1 | tail -2
timeout ${1:-45} python -m pytest tests -q -m gpu -x -p no:cacheprovider -v 2>&1 | grep -E "PASSED|FAILED|ERROR|passed|failed" | sed 's/ PASSED//' > gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
grep -c "::" gpurun_out/pytest_gpu_final.log
