#!/usr/bin/env bash
# Last single-GPU visit of round 2: all -m gpu tests, smoke, pairing probe, bench (full line), per-call distribution.
# Usage under gpurun: bash tools/gpu_final4.sh <tag>
set -u
TAG=${1:-final}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
echo "== pytest -m gpu" | tee $OUT/pytest_$TAG.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee -a $OUT/pytest_$TAG.log
echo "== smoke" | tee $OUT/smoke_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $OUT/smoke_$TAG.log
echo "== pairing probe"
timeout 60 python tools/pairing_probe.py 32 2>&1 | tee $OUT/pairing_probe_$TAG.log | head -12
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $OUT/bench_$TAG.log
cut -c1-500 $OUT/bench_$TAG.log
echo "== per-call distribution"
(timeout 200 python tools/e2e_dist.py 40 2>&1 | grep -v "^all") 2>&1 | tee $OUT/e2e_dist_$TAG.log
echo done
