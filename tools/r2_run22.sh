#!/bin/bash
# round 2, GPU call 22 (8 GPUs): tile invariance at 8 ranks with the captured step, then the default N=8 bench line (weak family + cfg4 extra)
mkdir -p gpurun_out
WCTB_CHECK_OUT=gpurun_out/r2_multi_gpu_check8_graph.json timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check8_graph.log 2>&1; echo "check exit $?"
grep multi_gpu_check gpurun_out/r2_multi_gpu_check8_graph.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_h2_8gpu_graph.json 2> gpurun_out/r2_bench_h2_8gpu_graph.err; echo "bench exit $?"
python - <<'P'
import json
for f in ('gpurun_out/r2_bench_h2_8gpu_graph.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d['config'].get('halo_exchanges_per_step'), d.get('cfg4'))
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace('.json','.err')).read()[-2500:])
P
grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench_h2_8gpu_graph.err | tail -5
