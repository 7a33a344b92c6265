#!/usr/bin/env bash
# ncu --set full captures (with source) of named kernels of the verify path.
# Usage under gpurun: bash tools/prof_kernels.sh <tag> <kernel-regex> [<kernel-regex> ...]
TAG=${1:-x}; shift
mkdir -p gpurun_out
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_$K python tools/prof_verify.py > gpurun_out/ncu_${TAG}_$K.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_$K.log | cut -c1-300
  ncu -i gpurun_out/prof_${TAG}_$K.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$K.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
