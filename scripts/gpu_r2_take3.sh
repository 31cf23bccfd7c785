#!/bin/bash
# Whole GPU suite on the per-variant shading tune (Perlin table in the small-scene kernels, draws taken together there and in the
# eight-wide kernels), then A/B against the code before it (variant orig).
set -u
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu16.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu16.log
tail -3 gpurun_out/pytest_gpu16.log
CFGS="c4 c1 c3 c5 scene3 scene5 c2" bash scripts/gpu_ab.sh orig 2>&1 | tail -2
CFGS="c4 c1 c3" bash scripts/gpu_ab.sh orig 2>&1 | tail -2
