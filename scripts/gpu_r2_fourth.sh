#!/bin/bash
# Round 2, fourth session: whole GPU suite on the current code, adaptive timings, the full default bench line (extras and CPU
# baseline included) and the reference arm, scene-creation stage times for 10^6 spheres.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for tp in default 32768; do
  if [ $tp = default ]; then timeout 300 python scripts/gpu_adaptive_times.py; else GS_TAIL_PIXELS=$tp timeout 300 python scripts/gpu_adaptive_times.py; fi
done 2>&1 | tee gpurun_out/adaptive_times.log
timeout 600 python scripts/gpu_reference_scenes.py > gpurun_out/reference_scenes.log 2>&1; tail -13 gpurun_out/reference_scenes.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cut -c1-300 gpurun_out/bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
for k in ('e2e','extra','cpu_baseline'): print(k, json.dumps(d.get(k))[:1500])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='issue'})
PY
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
GS_TRACE=1 timeout 600 python bench.py --config c5 --steps 2 --warmup 1 --e2e-steps 3 --no-extras --no-cpu-baseline 2> gpurun_out/b_c5.err | tail -1 > gpurun_out/c5.json
python -c "import json;d=json.load(open('gpurun_out/c5.json'));print(d['value'], d['e2e'])"
grep -E "gs_scene_create|build_nodes" gpurun_out/b_c5.err | tail -9
