#!/bin/bash
# Round 2, first contact: the new scene layout / eight-wide tree / stream discipline on a GPU.
# Quick correctness first (a failure here makes the rest pointless), then per-config numbers for both trees.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
# the old suite (binary tree for everything below 16384 primitives), then the round-2 file
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu > gpurun_out/pytest_old.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_old.log
tail -5 gpurun_out/pytest_old.log
timeout 1500 python -m pytest tests/test_gpu_round2.py -q -m gpu -s > gpurun_out/pytest_r2.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_r2.log
grep -E "passed|failed|FAILED|exit=|mismatches|RMSE" gpurun_out/pytest_r2.log | tail -40
# numbers: both trees on the configs where it could matter
: > gpurun_out/first.jsonl
for cfg in c4 c1 c3; do for w in 2 8; do
  timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --e2e-steps 2 --no-extras --no-cpu-baseline --bvh-width $w 2> gpurun_out/b_${cfg}_$w.err | tail -1 >> gpurun_out/first.jsonl
  tail -1 gpurun_out/first.jsonl | cut -c1-200
done; done
for w in 2 8; do
  GS_TRACE=1 timeout 600 python bench.py --config c5 --steps 3 --warmup 3 --e2e-steps 2 --no-extras --no-cpu-baseline --bvh-width $w 2> gpurun_out/b_c5_$w.err | tail -1 >> gpurun_out/first.jsonl
  tail -1 gpurun_out/first.jsonl | cut -c1-200
  grep -E "gs_scene_create|build_nodes" gpurun_out/b_c5_$w.err | tail -24
done
