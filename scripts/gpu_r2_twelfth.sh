#!/bin/bash
# Round 2, twelfth single-GPU session: work-queue chunk size in the adaptive rounds (and small fixed-count launches), with 1 and 2 chains.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "adaptive or multi or slices or fixed" > gpurun_out/pytest_gpu12.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu12.log
tail -3 gpurun_out/pytest_gpu12.log
: > gpurun_out/adaptive_chunk.log
for cfg in "256 1" "256 2" "128 2" "64 2" "32 2" "64 1" "64 3"; do
  set -- $cfg
  echo "GS_CHUNK=$1 GS_ADAPTIVE_CHAINS=$2" | tee -a gpurun_out/adaptive_chunk.log
  GS_CHUNK=$1 GS_ADAPTIVE_CHAINS=$2 timeout 300 python scripts/gpu_adaptive_times.py scene1 scene5 scene6 scene8 scene9 2>/dev/null | tee -a gpurun_out/adaptive_chunk.log
done
# fixed-count launches of 16 spp at 800x800 (10 M paths per launch) and the default bench step, chunk 256 vs 64
for c in 256 64 32; do
  echo "GS_CHUNK=$c: c4 16 spp per step" | tee -a gpurun_out/adaptive_chunk.log
  GS_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 --spp-per-step 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value']/1e6,1),'M paths/s', d['ms_per_step'])" | tee -a gpurun_out/adaptive_chunk.log
done
for c in 256 64; do
  echo "GS_CHUNK=$c: c4 500 spp per step" | tee -a gpurun_out/adaptive_chunk.log
  GS_CHUNK=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(round(d['value']/1e6,1),'M paths/s', d['ms_per_step'])" | tee -a gpurun_out/adaptive_chunk.log
done
