#!/bin/bash
# round 2, GPU call 12: does an A-operand start address off the 128-byte grid (the dx taps) slow tcgen05.mma?  + timing and
# ncu --set full of the generic h2 conv kernel at the cfg3 shapes
mkdir -p gpurun_out
timeout 300 python tools/h2_rates.py > gpurun_out/r2_h2_rates_off.txt 2>&1; tail -20 gpurun_out/r2_h2_rates_off.txt
timeout 300 python tools/profile_h2_generic.py > gpurun_out/r2_h2_generic_timing.txt 2>&1; cat gpurun_out/r2_h2_generic_timing.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_h2_kernel -c 1 -o gpurun_out/r2_h2_64 python tools/profile_h2_generic.py --once 64:64:540:960:0 > gpurun_out/r2_ncu_h2_64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_h2_kernel -c 1 -o gpurun_out/r2_h2_32p python tools/profile_h2_generic.py --once 32:32:1080:1920:1 > gpurun_out/r2_ncu_h2_32p.log 2>&1
ls -la gpurun_out/*.ncu-rep
