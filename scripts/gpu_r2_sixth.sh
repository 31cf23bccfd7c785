#!/bin/bash
# Round 2, sixth single-GPU session: whole GPU suite on the changed scene layout and wide-node decode, A/B of the PRMT-immediate
# node test and of a lower phase-B exit threshold for the eight-wide kernels, config-5 creation stages and bench line.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu6.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu6.log
tail -4 gpurun_out/pytest_gpu6.log
rm -f gpurun_out/variants.log
CFGS="c5 c4" bash scripts/gpu_ab.sh prmt0 msw16 msw20 2>&1 | tail -6
GS_TRACE=1 timeout 600 python bench.py --config c5 --steps 3 --warmup 2 --e2e-steps 3 --no-extras --no-cpu-baseline 2> gpurun_out/b_c5_6.err | tail -1 > gpurun_out/c5_6.json
grep -E "validate_desc|gs_scene_create|build_nodes" gpurun_out/b_c5_6.err | tail -14
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c5_6.json').read().strip().split('\n')[-1])
print('c5', round(d['value']/1e6,1), 'M paths/s; e2e', round(d['e2e']['value']/1e6,1), d['e2e'].get('stage_ms_rank0'))
PY
