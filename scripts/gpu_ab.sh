#!/bin/bash
# A/B session on the GPU box: fixed-seed images of the default build are the yardstick; every prebuilt variant
# (scripts/build_variants.sh) must reproduce them and is then timed.  One short python process per variant, each under
# its own timeout.  usage: bash scripts/gpu_ab.sh [variant names...]      (CFGS="c4 c1 c3")
mkdir -p gpurun_out
log=gpurun_out/variants.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader >> $log
GS_CUDA_LIB= timeout 120 python scripts/ab_bench.py save /tmp/ab_ref ${CFGS:-c4 c1 c3} 2>&1 | tee -a $log
names="$@"; [ -z "$names" ] && names="$(ls grayshift_b200/lib/variants 2>/dev/null)"
for v in $names; do
  skip=""; grep -q "GS_MEDIA_STREAM=0" grayshift_b200/lib/variants/$v/defs.txt 2>/dev/null && skip="c4 scene7"  # other random streams in media scenes
  GS_SKIP_CHECK="$skip" GS_VARIANT=$v GS_CUDA_LIB=$PWD/grayshift_b200/lib/variants/$v/libgrayshift_cuda.so timeout 90 python scripts/ab_bench.py check /tmp/ab_ref ${CFGS:-c4 c1 c3} 2>&1 | tail -3 | tee -a $log
  [ ${PIPESTATUS[0]} -ne 0 ] && echo "$v: FAILED or timed out" | tee -a $log
done
