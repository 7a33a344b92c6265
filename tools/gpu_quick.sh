#!/usr/bin/env bash
# Short GPU-box visit while iterating: parity tests, then probes.  Usage (under gpurun): bash tools/gpu_quick.sh [tag]
set -u
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest_$TAG.log
CKZG_B200_DEBUG=1 timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee -a $OUT/pytest_$TAG.log
echo "== probes" | tee $OUT/probe_$TAG.log
for M in 1 0; do
  CKZG_B200_STAGE1=$M PROBE_N=4096 timeout 300 python tools/gpu_probe.py modes 2>&1 | tail -2 | cut -c1-3000 | tee -a $OUT/probe_$TAG.log
done
PROBE_N=64 timeout 300 python tools/gpu_probe.py modes 2>&1 | tail -2 | cut -c1-3000 | tee -a $OUT/probe_$TAG.log
echo done
