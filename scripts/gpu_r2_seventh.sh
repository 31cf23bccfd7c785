#!/bin/bash
# Round 2, seventh single-GPU session: config-5 creation stages with the worker pool, config-5 bench line, kernel-by-kernel
# times of two small adaptive renders (where do scene1's 2.4 ms and scene3's 7 ms go?).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "wide or c5 or tree or multi or adaptive or create" > gpurun_out/pytest_gpu7.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu7.log
tail -3 gpurun_out/pytest_gpu7.log
GS_TRACE=1 timeout 600 python bench.py --config c5 --steps 3 --warmup 2 --e2e-steps 4 --no-extras --no-cpu-baseline 2> gpurun_out/b_c5_7.err | tail -1 > gpurun_out/c5_7.json
grep -E "validate_desc|gs_scene_create|build_nodes" gpurun_out/b_c5_7.err | tail -14
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c5_7.json').read().strip().split('\n')[-1])
print('c5', round(d['value']/1e6,1), 'M paths/s; e2e', round(d['e2e']['value']/1e6,1), d['e2e'].get('stage_ms_rank0'))
PY
A="python scripts/gpu_adaptive_times.py scene1 scene3"
timeout 300 $A > gpurun_out/adaptive_plain7.log 2>&1 && cat gpurun_out/adaptive_plain7.log && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/adaptive_launches7.csv $A > gpurun_out/adaptive_ncu7.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/adaptive_launches7.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
seq=[(r[ki].split('(')[0][-40:], float(r[vi].replace(',','')), r[ui]) for r in rows[1:]]
print(len(seq), 'launches')
for k,v,u in seq[:160]: print(f'{k:42s} {v:12.1f} {u}')
PY
