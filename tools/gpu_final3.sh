set -u
OUT=gpurun_out; mkdir -p $OUT
echo "== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > $OUT/bench_R3i.log
cut -c1-700 $OUT/bench_R3i.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_R3i.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > $OUT/ncu_bench_R3i.log 2>&1
tail -1 $OUT/ncu_bench_R3i.log | cut -c1-200
