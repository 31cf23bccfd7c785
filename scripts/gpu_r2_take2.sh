#!/bin/bash
# A/B in one session: the code before the RNG / Perlin-table changes (variant orig), single draws in the new structure (take0), and the default build
set -u
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
timeout 600 python -m pytest tests -q -m gpu -x -k "golden or seed or camera or media or beer" > gpurun_out/pytest_gpu15.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu15.log
tail -3 gpurun_out/pytest_gpu15.log
CFGS="c4 c1 c3 c5 scene3 scene5" bash scripts/gpu_ab.sh orig take0 2>&1 | tail -3
CFGS="c4 c1 c3" bash scripts/gpu_ab.sh orig take0 2>&1 | tail -3
