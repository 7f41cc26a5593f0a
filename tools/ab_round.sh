#!/bin/bash
# one short GPU slot: Gram kernel correctness + A/B timing, then the bench with each variant (no CPU baseline leg)
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -q -x -k "moments" 2>&1 | tail -5 | tee gpurun_out/gram_pytest.txt
timeout 40 python tools/gram_ab.py 2>&1 | tail -8
for v in 1 0; do
  WCTB_GRAM_VARIANT=$v timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gram_variant$v.json 2> gpurun_out/bench_gram_variant$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_gram_variant$v.json").read().strip().splitlines()[-1])
    print("variant $v: %.3f ms/step, %.1f MP/s, e2e %.1f MP/s" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print("variant $v: bench failed", e)
PY
done
