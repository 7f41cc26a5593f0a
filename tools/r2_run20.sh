#!/bin/bash
# round 2, GPU call 20 (2 GPUs): captured sharded step -- tile-invariance check (h2 captured, fp32 eager; one strip group), then the
# N=2 bench with the capture on, extras included (cfg5 = a second capture in the same process).  Tight timeouts: a hang must not eat the budget.
mkdir -p gpurun_out
WCTB_CHECK_OUT=gpurun_out/r2_multi_gpu_check2_graph.json timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check2_graph.log 2>&1; echo "check exit $?"
grep multi_gpu_check gpurun_out/r2_multi_gpu_check2_graph.log
WCTB_SHARD_GRAPH=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_h2_2gpu_graph.json 2> gpurun_out/r2_bench_h2_2gpu_graph.err; echo "bench(graph) exit $?"
python - <<'P'
import json
for f in ('gpurun_out/r2_bench_h2_2gpu_graph.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d['config'].get('halo_exchanges_per_step'), d.get('cfg5'))
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace('.json','.err')).read()[-2500:])
P
grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench_h2_2gpu_graph.err | tail -5
