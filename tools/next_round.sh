#!/bin/bash
# First GPU slot of the next round (about 90 s of box time): run everything that was written after round 1's last slot.
#   gpurun --timeout 240 -- 'bash tools/next_round.sh'
# 1. smoke + every test still marked pending_hw (strip-halo kernels, Gram ring variant 3, Newton-Schulz whitening)
# 2. Gram A/B including the peeled variant
# 3. whitening-solver A/B: critical-path timeline and bench line with WCTB_WHITEN=ns against the default
mkdir -p gpurun_out
timeout 120 python tools/hw_check.py 2>&1 | tail -25
timeout 40 python tools/gram_ab.py --peeled 2>&1 | tail -4
for solver in jacobi ns; do
  echo "--- stage timeline, WCTB_WHITEN=$solver" | tee -a gpurun_out/next_round_timeline.txt
  WCTB_WHITEN=$solver timeout 60 python tools/stage_timeline.py 2>&1 | tail -8 | tee -a gpurun_out/next_round_timeline.txt
  WCTB_WHITEN=$solver timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_whiten_$solver.json 2> gpurun_out/bench_whiten_$solver.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_whiten_$solver.json").read().strip().splitlines()[-1])
    print("WCTB_WHITEN=$solver: %.3f ms/step, %.1f MP/s, e2e %.1f MP/s" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("WCTB_WHITEN=$solver: bench failed", e)
PY
done
