#!/usr/bin/env bash
# Final single-GPU visit of a round: all -m gpu tests, smoke, bench (full line), ncu launch list of a short bench,
# targeted ncu --set full captures of the kernels that changed late (pairing, one-thread-per-butterfly FFT stage).
# Usage under gpurun: bash tools/gpu_final.sh <tag>
set -u
TAG=${1:-final}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
echo "== pytest -m gpu" | tee $OUT/pytest_$TAG.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee -a $OUT/pytest_$TAG.log
echo "== smoke" | tee $OUT/smoke_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $OUT/smoke_$TAG.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $OUT/bench_$TAG.log
cut -c1-400 $OUT/bench_$TAG.log
echo "== bench --impl reference (bounded: 2 steps)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.log | cut -c1-300
echo "== ncu launch list (short bench under ncu; numbers printed there are NOT bench values)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
tail -1 $OUT/ncu_bench_$TAG.log | cut -c1-200
echo "== ncu --set full: pairing (with source), one FFT stage with one thread per butterfly (2048 blobs)"
bash tools/prof_pairing_source.sh $TAG
timeout 900 ncu --set full --clock-control none -k regex:g1_fft_stage_thread_kernel -s 9 -c 1 -f -o /tmp/prof_${TAG}_fftthread python tools/prof_cells.py 2048 > $OUT/ncu_${TAG}_fftthread.log 2>&1
ncu -i /tmp/prof_${TAG}_fftthread.ncu-rep --page raw --csv > $OUT/raw_${TAG}_fftthread.csv 2>/dev/null
python tools/ncu_summary.py $OUT/raw_${TAG}_fftthread.csv > $OUT/ncu_${TAG}_fftthread_summary.txt 2>&1
head -20 $OUT/ncu_${TAG}_fftthread_summary.txt
echo done
