#!/bin/bash
# compute-sanitizer memcheck over small renders of every kernel family (scripts/gpu_sanitize.py), after a plain run of the same script.
set -u
mkdir -p gpurun_out
timeout 600 python scripts/gpu_sanitize.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -2 gpurun_out/sanitize_plain.log
timeout 2400 compute-sanitizer --tool ${TOOL:-memcheck} --error-exitcode 7 python scripts/gpu_sanitize.py > gpurun_out/sanitize_${TOOL:-memcheck}.log 2>&1; echo "sanitizer exit=$?" | tee -a gpurun_out/sanitize_${TOOL:-memcheck}.log
grep -E "ERROR SUMMARY|done|Invalid|Race|hazard" gpurun_out/sanitize_${TOOL:-memcheck}.log | head -10
