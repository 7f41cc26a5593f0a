#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_h2.py -q -x -k "fused or kd2sd or encoders" -s > gpurun_out/r2_pytest_fused.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_fused.log
grep -E "kd2sd|passed|failed|Error" gpurun_out/r2_pytest_fused.log | tail -8
timeout 300 python tools/profile_fused.py > gpurun_out/r2_fused_timing.txt 2>&1; cat gpurun_out/r2_fused_timing.txt
timeout 600 python tools/whiten_time.py > gpurun_out/r2_whiten_time.txt 2>&1; cat gpurun_out/r2_whiten_time.txt
WCTB_FAST_STATS_H2=1 timeout 600 python -m pytest tests/test_gpu_h2.py -q -s -k "cfg2 or cfg3 or five_stage" > gpurun_out/r2_pytest_h2_faststats.log 2>&1
grep -E "cfg2 .* h2|cfg3|h2 stage 1|passed|failed" gpurun_out/r2_pytest_h2_faststats.log
for e in 1e-4 1e-3 1e-2; do WCTB_FAST_STATS_H2=1 WCTB_EIG_EARLY=$e timeout 600 python -m pytest tests/test_gpu_h2.py -q -s -k "cfg2" 2>&1 | grep -E "cfg2 .* h2 " | sed "s/^/eig_early=$e /"; done
