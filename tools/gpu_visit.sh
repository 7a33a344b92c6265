#!/usr/bin/env bash
# Short GPU-box visit: parity tests, then the bench line with extras (no CPU baseline).  Usage (under gpurun):
#   bash tools/gpu_visit.sh [tag] [pytest -k expression]
set -u
TAG=${1:-v}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest_$TAG.log
if [ -n "$KEXPR" ]; then
  CKZG_B200_DEBUG=1 timeout 900 python -m pytest tests -q -m gpu -x -k "$KEXPR" 2>&1 | tail -25 | tee -a $OUT/pytest_$TAG.log
else
  CKZG_B200_DEBUG=1 timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee -a $OUT/pytest_$TAG.log
fi
echo "== bench" | tee $OUT/bench_$TAG.log
timeout 900 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -3 | tee -a $OUT/bench_$TAG.log
echo done
