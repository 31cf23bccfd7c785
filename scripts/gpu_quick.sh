#!/bin/bash
# quick A/B: smoke parity + short bench lines for c4 (headline), c1, c3, c5
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
PIPE=${PIPE:-megakernel}
for cfg in ${CFGS:-c4 c1 c3}; do
  python bench.py --config $cfg --pipeline $PIPE --steps 6 --warmup 2 --no-cpu-baseline --e2e-steps 1 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln.rstrip()); continue
    print('$PIPE $cfg', 'Mpaths/s %.1f' % (d['value']/1e6), 'Mrays/s %.1f' % d['mrays_per_s'], 'ms/step %.2f' % d['ms_per_step'], 'clk', d['clocks']['sm_mhz'])
"
done
