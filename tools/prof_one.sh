#!/usr/bin/env bash
# One ncu --set full capture (with source) of a kernel of the cells/proofs path, plus the probes.
# Usage under gpurun: bash tools/prof_one.sh <tag> <kernel-regex> <skip>
TAG=${1:-x}; K=${2:-g1_fft_stage}; S=${3:-20}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/prof_$TAG python tools/prof_cells.py 256 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
timeout 600 python tools/gpu_probe.py modes 2>&1 | tail -2 | cut -c1-2500 | tee gpurun_out/probe_$TAG.log
