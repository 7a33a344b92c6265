#!/usr/bin/env bash
# Final single-GPU visit of round 2 (second session): all -m gpu tests, smoke, bench (full line), reference arm,
# per-call distributions with the stream pool, pairing probe, ncu launch list of a short bench, ncu --set full of the pairing.
# Usage under gpurun: bash tools/gpu_final2.sh <tag>
set -u
TAG=${1:-final}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
echo "== pytest -m gpu" | tee $OUT/pytest_$TAG.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee -a $OUT/pytest_$TAG.log
echo "== smoke" | tee $OUT/smoke_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $OUT/smoke_$TAG.log
echo "== pairing probe"
timeout 60 python tools/pairing_probe.py 32 2>&1 | tee $OUT/pairing_probe_$TAG.log
echo "== per-call distributions (stream pool on), stage marks, concurrent callers"
(timeout 200 python tools/e2e_dist.py 40 2>&1 | grep -v "^all"
 timeout 150 python tools/e2e_probe.py 2>&1 | tail -1
 PROBE_N=64 timeout 150 python tools/e2e_dist.py 40 2>&1 | grep -v "^all"
 PROBE_CALLERS=2 timeout 150 python tools/e2e_probe.py 2>&1 | tail -1
 PROBE_CALLERS=3 timeout 150 python tools/e2e_probe.py 2>&1 | tail -1) 2>&1 | tee $OUT/e2e_pool_$TAG.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $OUT/bench_$TAG.log
cut -c1-600 $OUT/bench_$TAG.log
echo "== bench --impl reference (bounded: 2 steps)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.log | cut -c1-300
echo "== ncu launch list (short bench under ncu; numbers printed there are NOT bench values)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
tail -1 $OUT/ncu_bench_$TAG.log | cut -c1-200
echo "== ncu --set full: pairing (with source)"
bash tools/prof_pairing_source.sh $TAG
python tools/ncu_summary.py $OUT/raw_${TAG}_pairing.csv > $OUT/ncu_${TAG}_pairing_summary.txt 2>&1
cat $OUT/ncu_${TAG}_pairing_summary.txt | head -12
echo done
