#!/usr/bin/env bash
# ncu --set full of one multiplying G1-FFT stage over a whole 256-blob batch (CKZG_B200_FK_PARTS=1), and of
# the FK20 MSM kernel of the same batch.  Usage under gpurun: bash tools/prof_fft256.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
export CKZG_B200_FK_PARTS=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:g1_fft_stage_quad -s 18 -c 1 -f -o gpurun_out/prof_${TAG}_fft256 python tools/prof_cells.py 256 > gpurun_out/ncu_${TAG}_fft256.log 2>&1
tail -1 gpurun_out/ncu_${TAG}_fft256.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fk20_msm_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_msm256 python tools/prof_cells.py 256 > gpurun_out/ncu_${TAG}_msm256.log 2>&1
tail -1 gpurun_out/ncu_${TAG}_msm256.log | cut -c1-200
for k in fft256 msm256; do
  ncu -i gpurun_out/prof_${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$k.csv 2>/dev/null
done
