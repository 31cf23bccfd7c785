#!/bin/bash
# Round 2, second session: suite on the changed kernels (regeneration without the 64-bit division, in-kernel adaptive rounds,
# 256-bit loads), scene-creation stage times for 10^6 spheres, A/B of the prebuilt variants, ncu captures of C5 / C4.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu > gpurun_out/pytest_old.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_old.log
tail -4 gpurun_out/pytest_old.log
timeout 1500 python -m pytest tests/test_gpu_round2.py -q -m gpu -s > gpurun_out/pytest_r2.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_r2.log
grep -E "passed|failed|FAILED|exit=" gpurun_out/pytest_r2.log | tail -8
# adaptive scenes of main.rs against the CPU (launch-latency bound in round 1)
timeout 600 python scripts/gpu_reference_scenes.py > gpurun_out/reference_scenes.log 2>&1; tail -14 gpurun_out/reference_scenes.log
# 10^6 spheres: scene creation stage times + numbers for both trees
for w in 8 2; do
  GS_TRACE=1 timeout 600 python bench.py --config c5 --steps 3 --warmup 2 --e2e-steps 2 --no-extras --no-cpu-baseline --bvh-width $w 2> gpurun_out/b_c5_$w.err | tail -1 > gpurun_out/c5_$w.json
  cut -c1-160 gpurun_out/c5_$w.json; python -c "import json;d=json.load(open('gpurun_out/c5_$w.json'));print(d['e2e'])"
done
grep -E "gs_scene_create|build_nodes" gpurun_out/b_c5_8.err | tail -10
CFGS="c4 c1 c3 c5" bash scripts/gpu_ab.sh 2>&1 | tail -12
GS_BVH_WIDTH=2 GS_VARIANT=default_c5_binary python scripts/ab_bench.py check /tmp/ab_ref c5 2>&1 | tail -1 | tee -a gpurun_out/variants.log
GS_BVH_WIDTH=2 GS_VARIANT=ld128_c5_binary GS_CUDA_LIB=$PWD/grayshift_b200/lib/variants/ld128/libgrayshift_cuda.so python scripts/ab_bench.py check /tmp/ab_ref c5 2>&1 | tail -1 | tee -a gpurun_out/variants.log
# ncu: config 5 on the eight-wide tree, config 5 on the binary tree, config 4
C5="python bench.py --config c5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 --spp-per-step 2"
$C5 > gpurun_out/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 1 -c 1 -o gpurun_out/prof_c5_wide $C5 > gpurun_out/ncu_c5_wide.log 2>&1
$C5 --bvh-width 2 > gpurun_out/plain_c5b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 1 -c 1 -o gpurun_out/prof_c5_binary $C5 --bvh-width 2 > gpurun_out/ncu_c5_binary.log 2>&1
C4="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 --spp-per-step 16"
$C4 > gpurun_out/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 1 -c 1 -o gpurun_out/prof_c4 $C4 > gpurun_out/ncu_c4.log 2>&1
$C4 > gpurun_out/plain_c4b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $C4 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_c5_wide.log gpurun_out/ncu_c4.log
