#!/bin/bash
# ncu --set full capture of the render kernel on one config (default c5) -> gpurun_out/prof_<cfg>.ncu-rep
CFG=${1:-c5}; SPP=${2:-4}
mkdir -p gpurun_out
CMD="python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --spp-per-step $SPP"
$CMD > gpurun_out/plain_$CFG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 1 -c 1 -o gpurun_out/prof_$CFG $CMD > gpurun_out/ncu_$CFG.log 2>&1
tail -2 gpurun_out/ncu_$CFG.log; tail -1 gpurun_out/plain_$CFG.log | cut -c1-200
