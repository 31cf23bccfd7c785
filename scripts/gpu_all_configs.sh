#!/bin/bash
# bench.py on every BASELINE config (one JSON line each) -> gpurun_out/configs.jsonl
mkdir -p gpurun_out; : > gpurun_out/configs.jsonl
for cfg in ${CFGS:-c1 c2 c3 c4 c5}; do
  timeout 600 python bench.py --config $cfg --steps ${STEPS:-5} --warmup 3 --e2e-steps 2 ${EXTRA:-} 2> gpurun_out/bench_$cfg.err | tail -1 >> gpurun_out/configs.jsonl
  tail -1 gpurun_out/configs.jsonl | cut -c1-330
done
