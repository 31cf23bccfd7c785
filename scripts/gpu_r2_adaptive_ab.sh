#!/bin/bash
# A/B: adaptive batch rounds inside one cooperative launch (default) vs one launch per batch driven by the host (round-1 scheme)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -m gpu -k "adaptive" > gpurun_out/pytest_adaptive_kernel.log 2>&1; tail -2 gpurun_out/pytest_adaptive_kernel.log
GS_ADAPTIVE_ROUNDS=host timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -m gpu -k "adaptive" > gpurun_out/pytest_adaptive_host.log 2>&1; tail -2 gpurun_out/pytest_adaptive_host.log
for mode in kernel host kernel host; do
  echo "rounds=$mode"; GS_ADAPTIVE_ROUNDS=$mode timeout 300 python scripts/gpu_adaptive_times.py scene1 scene3 scene5 scene6 scene8 scene9 scene11
done 2>&1 | tee gpurun_out/adaptive_ab.log
