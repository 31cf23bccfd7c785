#!/bin/bash
# shared-memory carve-out hint of the render kernels (GS_CARVEOUT, percent of the maximum): default build, run-time knob
mkdir -p gpurun_out
log=gpurun_out/carveout.log
GS_CUDA_LIB= timeout 120 python scripts/ab_bench.py save /tmp/ab_ref ${CFGS:-c4 c1 c3 c5} 2>&1 | tee -a $log
for c in ${SWEEP:-0 5 10 15 25 50 100}; do
  GS_CARVEOUT=$c GS_VARIANT=carveout$c timeout 120 python scripts/ab_bench.py check /tmp/ab_ref ${CFGS:-c4 c1 c3 c5} 2>&1 | tail -1 | tee -a $log
done
