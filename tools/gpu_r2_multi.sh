#!/usr/bin/env bash
# Round-2 multi-GPU visit (gpurun --gpus N): the -m gpu tests that span devices, then bench.py under torchrun.
# Usage: bash tools/gpu_r2_multi.sh <tag> <N> [bench args...]
set -u
TAG=${1:-R2m}; N=${2:-2}; shift; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,memory.used --format=csv > $OUT/gpu_$TAG.txt 2>&1
nvidia-smi topo -m >> $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt
echo "== pytest (multi-device tests)" | tee $OUT/pytest_$TAG.log
CKZG_B200_DEBUG=1 timeout 900 python -m pytest tests -q -m gpu -x -k "multi or consumer or sharded" 2>&1 | tail -15 | tee -a $OUT/pytest_$TAG.log
echo "== bench N=$N $*" | tee $OUT/bench_$TAG.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" 2>&1 | tail -6 | tee -a $OUT/bench_$TAG.log
echo done
