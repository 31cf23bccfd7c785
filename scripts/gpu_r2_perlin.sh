#!/bin/bash
# A/B: Perlin gradient from the CTA's table (default build) against the select-based form (variant perlin0)
set -u
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
CFGS="c4 scene3 scene5 c1" bash scripts/gpu_ab.sh perlin0 2>&1 | tail -4
CFGS="c4 scene3" bash scripts/gpu_ab.sh perlin0 2>&1 | tail -2
