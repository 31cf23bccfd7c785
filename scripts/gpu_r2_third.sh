#!/bin/bash
# Round 2, third session: adaptive rounds with the round state in shared memory, centre/half-extent binary nodes, A/B of
# shared-memory stacks, occupancy and 256-bit loads (config 5 on both trees), scene-creation stage times.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -m gpu -k "adaptive or tiny or 2_32 or slices or reproducible or golden or same_seed or closest_hit_ids or both_trees" > gpurun_out/pytest_subset.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_subset.log
tail -4 gpurun_out/pytest_subset.log
for tp in default 256 8192; do
  if [ $tp = default ]; then timeout 300 python scripts/gpu_adaptive_times.py; else GS_TAIL_PIXELS=$tp timeout 300 python scripts/gpu_adaptive_times.py; fi
done 2>&1 | tee gpurun_out/adaptive_times.log
GS_TRACE=1 timeout 600 python bench.py --config c5 --steps 3 --warmup 2 --e2e-steps 3 --no-extras --no-cpu-baseline 2> gpurun_out/b_c5.err | tail -1 > gpurun_out/c5.json
cut -c1-160 gpurun_out/c5.json; python -c "import json;d=json.load(open('gpurun_out/c5.json'));print(d['e2e'])"
grep -E "gs_scene_create|build_nodes" gpurun_out/b_c5.err | tail -9
CFGS="c4 c1 c3 c5" bash scripts/gpu_ab.sh 2>&1 | tail -12
for v in default ld256all smem8 blocks6 blocks7 blocks8 blocks8smem8; do
  lib=""; [ $v != default ] && lib=$PWD/grayshift_b200/lib/variants/$v/libgrayshift_cuda.so
  GS_BVH_WIDTH=2 GS_VARIANT=${v}_c5_binary GS_CUDA_LIB=$lib timeout 200 python scripts/ab_bench.py check /tmp/ab_ref c5 2>&1 | tail -1 | tee -a gpurun_out/variants.log
done
