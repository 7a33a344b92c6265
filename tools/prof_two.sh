#!/usr/bin/env bash
# Two ncu --set full captures (with source) from one cells/proofs run: the FK20 MSM kernel and one
# multiplying G1-FFT stage; raw pages exported for tools/ncu_summary.py.
# Usage under gpurun: bash tools/prof_two.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk20_msm_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_msm python tools/prof_cells.py 256 > gpurun_out/ncu_${TAG}_msm.log 2>&1
tail -1 gpurun_out/ncu_${TAG}_msm.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:g1_fft_stage_quad -s 18 -c 1 -f -o gpurun_out/prof_${TAG}_fft python tools/prof_cells.py 256 > gpurun_out/ncu_${TAG}_fft.log 2>&1
tail -1 gpurun_out/ncu_${TAG}_fft.log | cut -c1-200
for k in msm fft; do
  ncu -i gpurun_out/prof_${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$k.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
