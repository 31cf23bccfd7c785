#!/bin/bash
for cfg in ${CFGS:-c4 c1 c3 c5}; do
 for leaf in ${LEAVES:-1 2 4 8}; do
  python bench.py --config $cfg --max-leaf $leaf --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    print('$cfg leaf<=$leaf Mpaths/s %.1f Mrays/s %.1f flatten_s %s' % (d['value']/1e6, d['mrays_per_s'], d['config']['host_flatten_s']))
"
 done
done
