#!/bin/bash
# Bench every prebuilt variant under grayshift_b200/lib/variants/ (plus the default build) on the GPU box.
# usage: bash scripts/gpu_variants.sh [variant names...]      (CFGS="c4 c1 c3", STEPS=4)
mkdir -p gpurun_out
names="$@"; [ -z "$names" ] && names="default $(ls grayshift_b200/lib/variants 2>/dev/null)"
for v in $names; do
  lib=""; [ "$v" != default ] && lib=$PWD/grayshift_b200/lib/variants/$v/libgrayshift_cuda.so
  for cfg in ${CFGS:-c4 c1 c3}; do
    GS_CUDA_LIB=$lib timeout 300 python bench.py --config $cfg --steps ${STEPS:-4} --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print('%-22s %s Mpaths/s %8.1f Mrays/s %8.1f' % ('$v', '$cfg', d['value']/1e6, d['mrays_per_s']))
" | tee -a gpurun_out/variants.log
  done
done
