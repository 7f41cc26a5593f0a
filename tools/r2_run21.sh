#!/bin/bash
# round 2, GPU call 21 (4 GPUs): captured sharded step with interior ranks (two neighbours) + clean exit (parallel.shutdown)
mkdir -p gpurun_out
WCTB_CHECK_OUT=gpurun_out/r2_multi_gpu_check4_graph.json timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check4_graph.log 2>&1; echo "check exit $?"
grep multi_gpu_check gpurun_out/r2_multi_gpu_check4_graph.log
WCTB_SHARD_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_h2_4gpu_graph.json 2> gpurun_out/r2_bench_h2_4gpu_graph.err; echo "bench(graph) exit $?"
python - <<'P'
import json
for f in ('gpurun_out/r2_bench_h2_4gpu_graph.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d['config'].get('halo_exchanges_per_step'))
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace('.json','.err')).read()[-2500:])
P
grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench_h2_4gpu_graph.err | tail -5
