#!/bin/bash
# round 2, GPU call 1: first contact of the h2 engine + the round-1 kernels that never ran on hardware
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_run1_gpu.txt 2>&1
timeout 600 python tools/h2_check.py --time > gpurun_out/r2_h2_check.log 2>&1; echo "h2_check exit $?" >> gpurun_out/r2_h2_check.log
timeout 900 python -m pytest tests/test_gpu_h2.py -x -q -s > gpurun_out/r2_pytest_h2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_h2.log
WCTB_PENDING_HW=1 timeout 900 python -m pytest tests -q -m "gpu and pending_hw" > gpurun_out/r2_pytest_pending_hw.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_pending_hw.log
tail -5 gpurun_out/r2_h2_check.log; tail -15 gpurun_out/r2_pytest_h2.log; tail -15 gpurun_out/r2_pytest_pending_hw.log
