#!/bin/bash
# Round 2, final single-GPU session: whole GPU suite + smoke on the final code, the default bench line and the reference arm,
# tree-choice threshold, adaptive tail threshold with host-driven rounds, scene-creation stages, ncu evidence of the final kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
for tp in 2048 8192; do GS_TAIL_PIXELS=$tp timeout 300 python scripts/gpu_adaptive_times.py; done 2>&1 | tee gpurun_out/adaptive_times.log
timeout 600 python scripts/gpu_tree_threshold.py 2>&1 | tee gpurun_out/tree_threshold.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cut -c1-200 gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-160 gpurun_out/bench_ref.json
EXTRA="--no-extras" bash scripts/gpu_all_configs.sh > gpurun_out/all_configs.log 2>&1; tail -5 gpurun_out/all_configs.log
CFGS="c4 c1 c3" bash scripts/gpu_ab.sh 2>&1 | tail -6
# the wavefront alternative with and without its material-sorted shading queues
: > gpurun_out/wavefront.jsonl
for cfg in c4 c1 c3; do for srt in 1 0; do
  timeout 300 python bench.py --config $cfg --steps 3 --warmup 2 --e2e-steps 1 --no-extras --no-cpu-baseline --pipeline wavefront --wavefront-sort $srt 2>/dev/null | tail -1 >> gpurun_out/wavefront.jsonl
  python -c "import json;d=json.loads(open('gpurun_out/wavefront.jsonl').read().strip().split('\n')[-1]);print('wavefront $cfg sort=$srt', round(d['value']/1e6,1), 'M paths/s')"
done; done
GS_TRACE=1 timeout 600 python bench.py --config c5 --steps 2 --warmup 1 --e2e-steps 3 --no-extras --no-cpu-baseline 2> gpurun_out/b_c5.err | tail -1 > gpurun_out/c5.json
grep -E "gs_scene_create|build_nodes" gpurun_out/b_c5.err | tail -13
# ncu: launch list of a short default run, full captures of the dominant kernel on C4 and C5
C4="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 --spp-per-step 16"
$C4 > gpurun_out/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $C4 > gpurun_out/ncu_launches.log 2>&1
$C4 > gpurun_out/plain_c4b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 1 -c 1 -o gpurun_out/prof_c4 $C4 > gpurun_out/ncu_c4.log 2>&1
C5="python bench.py --config c5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 --spp-per-step 2"
$C5 > gpurun_out/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 1 -c 1 -o gpurun_out/prof_c5 $C5 > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c4.log gpurun_out/ncu_c5.log
