#!/bin/bash
# Round 2, last single-GPU session on the final code: whole GPU suite + smoke, the default bench line and the reference arm,
# ncu launch list of the bench command, full captures of the config-4 and config-5 kernels, every config.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cut -c1-200 gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-160 gpurun_out/bench_ref.json
C4="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 --spp-per-step 16"
$C4 > gpurun_out/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $C4 > gpurun_out/ncu_launches.log 2>&1
tail -n 3 gpurun_out/launches_c4.csv | cut -c1-200
$C4 > gpurun_out/plain_c4b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 1 -c 1 -o gpurun_out/prof_c4 $C4 > gpurun_out/ncu_c4.log 2>&1
C5="python bench.py --config c5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 --spp-per-step 2"
$C5 > gpurun_out/plain_c5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 1 -c 1 -o gpurun_out/prof_c5 $C5 > gpurun_out/ncu_c5.log 2>&1
tail -n 1 gpurun_out/ncu_c4.log; tail -n 1 gpurun_out/ncu_c5.log
CFGS="c1 c2 c3 c4" EXTRA="--no-extras --no-cpu-baseline" bash scripts/gpu_all_configs.sh > gpurun_out/all_configs.log 2>&1; tail -2 gpurun_out/all_configs.log | cut -c1-120
