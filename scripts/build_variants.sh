#!/bin/bash
# Build differently tuned copies of libgrayshift_cuda.so HERE (nvcc cross-compiles, no GPU time) into
# grayshift_b200/lib/variants/<name>/ ; they travel to the GPU box with the snapshot and are selected with GS_CUDA_LIB.
# usage: bash scripts/build_variants.sh name1 "<defs 1>" name2 "<defs 2>" ...
set -u
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  out=../lib/variants/$name
  mkdir -p grayshift_b200/lib/variants/$name
  if make -C grayshift_b200/csrc -j4 OUT=$out NVDEFS="$defs" > grayshift_b200/lib/variants/$name/build.log 2>&1; then
    rm -f grayshift_b200/lib/variants/$name/*.o
    echo "$name [$defs]: $(grep -A2 'render_fixed_kernelILi0' grayshift_b200/lib/variants/$name/render_kernels.ptxas.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores|[0-9]+ bytes stack frame' | tr '\n' ' ')"
    echo "$defs" > grayshift_b200/lib/variants/$name/defs.txt
  else
    echo "$name BUILD FAILED"; tail -5 grayshift_b200/lib/variants/$name/build.log
  fi
done
