#!/bin/bash
# multi-GPU bench exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit=$?" >> gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_same_box.json 2>/dev/null
tail -3 gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_n$N.json; cut -c1-200 gpurun_out/bench_n1_same_box.json
