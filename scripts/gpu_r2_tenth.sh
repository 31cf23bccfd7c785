#!/bin/bash
# Round 2, tenth single-GPU session: ncu capture of the adaptive megakernel variant on the final scene at 400x400 (first round of
# the second render), to set beside the fixed-count kernel's capture (profiles/r02_v5_c4_*).
set -u
mkdir -p gpurun_out
A="python scripts/gpu_adaptive_times.py scene8"
timeout 300 $A > gpurun_out/adaptive_plain10.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_fixed -s 34 -c 1 -o gpurun_out/prof_scene8_adaptive $A > gpurun_out/ncu_scene8.log 2>&1
tail -2 gpurun_out/ncu_scene8.log
