#!/bin/bash
# round 2, GPU call 16 (1 GPU): is the generic kernel bound by DRAM access or inside the SM?  same layer, L2-resident vs flushed, three sizes
mkdir -p gpurun_out
S="64:64:270:480:0 64:64:540:960:0 64:64:1080:1920:0 32:32:540:960:0 32:32:1080:1920:0 16:32:1080:1920:0 32:16:1080:1920:0 128:128:270:480:0"
(timeout 300 python tools/profile_h2_generic.py $S; timeout 300 python tools/profile_h2_generic.py --noflush $S; echo "--- streamed"; WCTB_H2_RESIDENT=0 timeout 300 python tools/profile_h2_generic.py --noflush 64:64:270:480:0 64:64:540:960:0 32:32:540:960:0) 2>&1 | tee gpurun_out/r2_h2_generic_l2.txt
