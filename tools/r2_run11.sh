#!/bin/bash
# round 2, GPU call 11 (8 GPUs): tile invariance at 8 ranks, N=4 and N=8 bench (N=8 line carries cfg4 = BASELINE configs[3])
mkdir -p gpurun_out
nvidia-smi -L | wc -l
WCTB_CHECK_OUT=gpurun_out/r2_multi_gpu_check8.json timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check8.log 2>&1; echo "check8 exit $?"; grep multi_gpu_check gpurun_out/r2_multi_gpu_check8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_h2_4gpu.json 2> gpurun_out/r2_bench_h2_4gpu.err; echo "bench4 exit $?"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_h2_8gpu.json 2> gpurun_out/r2_bench_h2_8gpu.err; echo "bench8 exit $?"
python - <<'P'
import json
for f in ('gpurun_out/r2_bench_h2_4gpu.json','gpurun_out/r2_bench_h2_8gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d['config'].get('halo_exchanges_per_step'), d.get('cfg4'))
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace('.json','.err')).read()[-2500:])
P
