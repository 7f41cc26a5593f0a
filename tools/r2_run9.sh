#!/bin/bash
# round 2, GPU call 9 (2 GPUs): tile-invariance pytest, N=2 bench (weak family + cfg5 extra)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -q -s > gpurun_out/r2_pytest_multi_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_multi_gpu.log
grep -E "multi_gpu_check|passed|failed|Error|error" gpurun_out/r2_pytest_multi_gpu.log | tail -12
WCTB_PEER_HALO=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_h2_2gpu_nopeer.json 2> gpurun_out/r2_bench_h2_2gpu_nopeer.err; echo "bench(no peer) exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_h2_2gpu_nopeer.json').read().strip().splitlines()[-1]); print('no-peer', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_h2_2gpu.json 2> gpurun_out/r2_bench_h2_2gpu.err; echo "bench exit $?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_h2_2gpu.json').read().strip().splitlines()[-1])
    print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d.get('cfg5'))
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/r2_bench_h2_2gpu.err').read()[-3000:])
P
