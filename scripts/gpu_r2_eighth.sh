#!/bin/bash
# Round 2, eighth single-GPU session: the adaptive tail that renders batches ahead -- parity tests, then wall times of the
# reference's adaptive scenes under different pixel thresholds for the 4- and 8-warp-per-pixel kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "adaptive or multi" > gpurun_out/pytest_gpu8.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu8.log
tail -3 gpurun_out/pytest_gpu8.log
: > gpurun_out/tail_ahead.log
for cfg in "0 0" "2048 512" "8192 512" "8192 2048" "0 8192" "1024 256"; do
  set -- $cfg
  echo "GS_TAIL_AHEAD4=$1 GS_TAIL_AHEAD8=$2" | tee -a gpurun_out/tail_ahead.log
  GS_TRACE=1 GS_TAIL_AHEAD4=$1 GS_TAIL_AHEAD8=$2 timeout 300 python scripts/gpu_adaptive_times.py 2> gpurun_out/tail_ahead_trace_$1_$2.err | tee -a gpurun_out/tail_ahead.log
done
grep "pixels left" gpurun_out/tail_ahead_trace_0_0.err | sort | uniq -c | sort -rn | head -12
