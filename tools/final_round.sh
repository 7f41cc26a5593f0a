#!/bin/bash
# last GPU slot of the round: Gram A/B (3 variants), official-shape bench line (no CPU leg) with the default and with the
# previous Gram kernel, I/O row timing, then as much of the GPU suite as fits
mkdir -p gpurun_out
timeout 30 python tools/gram_ab.py 2>&1 | tail -4
timeout 60 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
WCTB_GRAM_VARIANT=2 timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_gram2.json 2> gpurun_out/bench_final_gram2.err
python - <<'PY'
import json
for n in ("bench_final", "bench_final_gram2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print("%s: %.3f ms/step, %.1f MP/s, e2e %.1f MP/s, clocks %s" % (n, d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("clocks")))
    except Exception as e:
        print(n, "failed", e)
PY
timeout 25 python tools/io_bench.py 2>&1 | tail -1 | cut -c1-900
timeout ${1:-45} python -m pytest tests/test_gpu_parity.py tests/test_image_io_gpu.py tests/test_gpu_umma.py -q -m gpu -x -p no:cacheprovider -v 2>&1 | grep -E "PASSED|FAILED|ERROR|passed|failed" | sed 's/ PASSED//' > gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
grep -c "::" gpurun_out/pytest_gpu_final.log
