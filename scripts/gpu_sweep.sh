#!/bin/bash
# A/B sweep of compile-time knobs ON the GPU box (it has nvcc): each variant = rebuild the CUDA library + short benches.
# usage: bash scripts/gpu_sweep.sh "<defs variant 1>" "<defs variant 2>" ...     (CFGS="c4 c3" to choose configs)
for defs in "$@"; do
  make -C grayshift_b200/csrc clean > /dev/null 2>&1
  make -C grayshift_b200/csrc -j3 NVDEFS="$defs" > /tmp/build.log 2>&1 || { echo "BUILD FAILED: $defs"; tail -5 /tmp/build.log; continue; }
  make -C grayshift_b200/host > /dev/null 2>&1
  regs=$(grep -A2 "render_fixed_kernelILb0" grayshift_b200/lib/render_kernels.ptxas.log | grep -oE "Used [0-9]+ registers|[0-9]+ bytes spill stores" | tr '\n' ' ')
  for cfg in ${CFGS:-c4 c1 c3}; do
    python bench.py --config $cfg --pipeline ${PIPE:-megakernel} --steps ${STEPS:-4} --warmup 2 --no-cpu-baseline --e2e-steps 1 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print('[$defs] [$regs] $cfg Mpaths/s %.1f Mrays/s %.1f' % (d['value']/1e6, d['mrays_per_s']))
"
  done
done
# leave the default build behind
make -C grayshift_b200/csrc clean > /dev/null 2>&1; make -C grayshift_b200/csrc -j3 > /dev/null 2>&1; make -C grayshift_b200/host > /dev/null 2>&1
