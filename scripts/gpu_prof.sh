#!/bin/bash
# ncu: launch list + one full capture of a kernel (regex in $1, default wf_traverse) on a short C4 run
K=${1:-wf_traverse}
SKIP=${2:-6}
mkdir -p gpurun_out
NCU_CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --spp-per-step 8"
$NCU_CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
$NCU_CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o gpurun_out/prof_$K $NCU_CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
