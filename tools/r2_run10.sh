#!/bin/bash
# round 2, GPU call 10 (2 GPUs): N=2 bench, peer-memory halo vs NCCL exchange
mkdir -p gpurun_out
WCTB_PEER_HALO=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_h2_2gpu_nopeer.json 2> gpurun_out/r2_bench_h2_2gpu_nopeer.err; echo "bench(no peer) exit $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_h2_2gpu.json 2> gpurun_out/r2_bench_h2_2gpu.err; echo "bench exit $?"
python - <<'P'
import json
for f in ('gpurun_out/r2_bench_h2_2gpu_nopeer.json','gpurun_out/r2_bench_h2_2gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d.get('cfg5'))
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace('.json','.err')).read()[-1500:])
P
