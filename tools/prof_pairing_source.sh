#!/usr/bin/env bash
# ncu --set full with source correlation of ONE pairing_check_kernel launch (n = 64 verify call); exports the SASS-level
# page (stall samples per instruction) for tools/ncu_source_top.py.  Usage under gpurun: bash tools/prof_pairing_source.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairing_check_kernel -s 2 -c 1 -f -o /tmp/prof_${TAG}_pairing \
    python tools/prof_verify.py 64 > gpurun_out/ncu_${TAG}_pairing.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_pairing.log | cut -c1-200
ncu -i /tmp/prof_${TAG}_pairing.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_pairing.csv 2>/dev/null
ncu -i /tmp/prof_${TAG}_pairing.ncu-rep --page source --csv --print-source sass > /tmp/src_${TAG}_pairing.csv 2>/dev/null
python tools/ncu_source_top.py /tmp/src_${TAG}_pairing.csv > gpurun_out/ncu_${TAG}_pairing_source_top.txt 2>&1
head -c 3000 gpurun_out/ncu_${TAG}_pairing_source_top.txt
xz -9 -c /tmp/src_${TAG}_pairing.csv > gpurun_out/src_${TAG}_pairing.csv.xz
ls -la gpurun_out/src_${TAG}_pairing.csv.xz
