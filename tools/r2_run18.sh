#!/bin/bash
# round 2, GPU call 18 (2 GPUs): sharded step captured in a CUDA graph (tile invariance incl. replay vs eager), N=2 bench with the
# graph and with WCTB_SHARD_GRAPH=0 (eager schedule) for the A/B
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m pytest tests/test_multi_gpu.py -q -x -s > gpurun_out/r2_pytest_multi_gpu_graph.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_multi_gpu_graph.log
grep -E "multi_gpu_check|passed|failed|exit|Error|error" gpurun_out/r2_pytest_multi_gpu_graph.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_h2_2gpu_graph.json 2> gpurun_out/r2_bench_h2_2gpu_graph.err; echo "bench(graph) exit $?"
WCTB_SHARD_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_h2_2gpu_eager.json 2> gpurun_out/r2_bench_h2_2gpu_eager.err; echo "bench(eager) exit $?"
python - <<'P'
import json
for f in ('gpurun_out/r2_bench_h2_2gpu_graph.json','gpurun_out/r2_bench_h2_2gpu_eager.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d['config'].get('halo_exchanges_per_step'))
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace('.json','.err')).read()[-2500:])
P
tail -3 gpurun_out/r2_bench_h2_2gpu_graph.err
