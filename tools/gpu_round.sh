#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench, launch list (ncu), optional full capture.
# Usage (under gpurun): bash tools/gpu_round.sh [tag] [full-capture-kernel-regex]
set -u
TAG=${1:-r01}
KREGEX=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu" | tee $OUT/pytest_$TAG.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 | tee -a $OUT/pytest_$TAG.log
echo "== smoke" | tee $OUT/smoke_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a $OUT/smoke_$TAG.log
echo "== bench" | tee $OUT/bench_$TAG.log
timeout 900 python bench.py 2>&1 | tail -4 | tee -a $OUT/bench_$TAG.log
echo "== ncu launch list (short bench under ncu; numbers printed there are NOT bench values)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --blobs 1024 --no-extra --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
tail -2 $OUT/ncu_bench_$TAG.log
if [ -n "$KREGEX" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 1 -c 1 -o $OUT/prof_$TAG \
      python bench.py --steps 1 --warmup 1 --blobs 512 --no-extra --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
  tail -2 $OUT/ncu_full_$TAG.log
fi
echo "== probes" | tee $OUT/probe_$TAG.log
timeout 600 python tools/gpu_probe.py 2>&1 | tail -8 | cut -c1-1500 | tee -a $OUT/probe_$TAG.log
echo done
