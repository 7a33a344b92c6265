#!/usr/bin/env bash
# Round-2 GPU-box visit: parity tests (all -m gpu), smoke, optional bench.  Usage: bash tools/gpu_r2.sh <tag> [pytest -k expr] [bench args...]
set -u
TAG=${1:-R2}; KEXPR=${2:-}; shift; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,memory.used --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; free -g | head -2 >> $OUT/gpu_$TAG.txt
echo "== pytest -m gpu" | tee $OUT/pytest_$TAG.log
if [ -n "$KEXPR" ]; then
  CKZG_B200_DEBUG=1 timeout 1500 python -m pytest tests -q -m gpu -x -k "$KEXPR" --durations=8 2>&1 | tail -40 | tee -a $OUT/pytest_$TAG.log
else
  CKZG_B200_DEBUG=1 timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 2>&1 | tail -40 | tee -a $OUT/pytest_$TAG.log
fi
if [ $# -gt 0 ]; then
  echo "== bench $*" | tee $OUT/bench_$TAG.log
  timeout 900 python bench.py "$@" 2>&1 | tail -3 | tee -a $OUT/bench_$TAG.log
fi
echo done
