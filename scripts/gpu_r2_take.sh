#!/bin/bash
# Parity on the draws-taken-together RNG refactor and the small-scene Perlin table, then A/B against single draws (variant take0)
set -u
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu14.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu14.log
tail -3 gpurun_out/pytest_gpu14.log
CFGS="c4 c1 c3 c5 scene3 scene5" bash scripts/gpu_ab.sh take0 2>&1 | tail -3
CFGS="c4 c1 c3" bash scripts/gpu_ab.sh take0 2>&1 | tail -2
