#!/bin/bash
# Round 2, ninth single-GPU session: adaptive parity tests on the final tail code, wall times of the reference's adaptive scenes,
# and the kernel-by-kernel times of the final scene at 400x400 (scene8): where do its rounds lose against the fixed-count rate?
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "adaptive or multi" > gpurun_out/pytest_gpu9.log 2>&1; echo "exit=$?" >> gpurun_out/pytest_gpu9.log
tail -3 gpurun_out/pytest_gpu9.log
GS_TRACE=1 timeout 300 python scripts/gpu_adaptive_times.py 2> gpurun_out/adaptive9_trace.err | tee gpurun_out/adaptive9.log
grep "pixels left" gpurun_out/adaptive9_trace.err | uniq -c | head -60
A="python scripts/gpu_adaptive_times.py scene8"
timeout 300 $A > gpurun_out/adaptive_plain9.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/adaptive_launches9.csv $A > gpurun_out/adaptive_ncu9.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/adaptive_launches9.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki].split('(')[0][-34:], float(r[vi].replace(',',''))) for r in rows[1:]]
n=len(seq)//4
print(len(seq),'launches; last render:')
for k,v in seq[-n:]: print(f'{k:36s} {v/1e3:10.1f} us')
PY
