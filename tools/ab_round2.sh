#!/bin/bash
# one short GPU slot: first-layer + eigensolver early-stop A/B (tests, stage timeline, bench)
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -q -x -k "conv_first or eigh or whiten or five_stage or two_stream or style_cache or truncation" 2>&1 | tail -5 | tee gpurun_out/ab2_pytest.txt
echo "--- timeline: old first kernel, early 3e-6" | tee gpurun_out/ab2_timeline.txt
WCTB_FIRST_VARIANT=1 WCTB_EIG_EARLY=3e-6 timeout 60 python tools/stage_timeline.py 2>&1 | tail -8 | tee -a gpurun_out/ab2_timeline.txt
echo "--- timeline: new defaults" | tee -a gpurun_out/ab2_timeline.txt
timeout 60 python tools/stage_timeline.py 2>&1 | tail -8 | tee -a gpurun_out/ab2_timeline.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    print("$name: %.3f ms/step, %.1f MP/s, e2e %.1f MP/s" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$name: bench failed", e)
PY
}
run base WCTB_FIRST_VARIANT=1 WCTB_EIG_EARLY=3e-6
run first2 WCTB_EIG_EARLY=3e-6
run defaults WCTB_X=0
