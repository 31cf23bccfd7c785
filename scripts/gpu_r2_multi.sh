#!/bin/bash
# Round 2, multi-GPU session (gpurun --gpus N): gs_render_multi tests, the torchrun bench with its extras (config 5, strong
# scaling, single-process gs_render_multi), and gs_render_multi with the NCCL transport.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_$N.txt
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "multi" -s > gpurun_out/pytest_multi_$N.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_multi_$N.log
tail -4 gpurun_out/pytest_multi_$N.log
GS_MULTI_REDUCE=nccl timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "multi" -s > gpurun_out/pytest_multi_nccl_$N.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_multi_nccl_$N.log
tail -4 gpurun_out/pytest_multi_nccl_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
print(json.dumps(d.get('extra'), indent=1)[:3000])
PY
tail -5 gpurun_out/bench_n$N.err
