#!/bin/bash
# Round 2, eleventh single-GPU session: chains of adaptive rounds on several streams -- parity tests, then wall times for 1..4 chains.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "adaptive or multi" > gpurun_out/pytest_gpu11.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu11.log
tail -3 gpurun_out/pytest_gpu11.log
: > gpurun_out/adaptive_chains.log
for k in 1 2 3 4; do
  echo "GS_ADAPTIVE_CHAINS=$k" | tee -a gpurun_out/adaptive_chains.log
  GS_ADAPTIVE_CHAINS=$k timeout 300 python scripts/gpu_adaptive_times.py scene1 scene3 scene5 scene6 scene8 scene9 scene11 2>/dev/null | tee -a gpurun_out/adaptive_chains.log
done
