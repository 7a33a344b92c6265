#!/usr/bin/env bash
# Round-2 late visit: radix-29 wide product A/B, stream pool A/B.  Usage (under gpurun): bash tools/gpu_r3.sh [tag]
set -u
TAG=${1:-R3e}
OUT=gpurun_out
mkdir -p $OUT
(timeout 300 python -m pytest tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -3) | tee $OUT/pytest_$TAG.log
(echo "== radix-29 product (default build)"; timeout 60 python tools/pairing_probe.py 32 2>&1
 if [ -f c-kzg-4844_b200/build/libckzg_b200_w32.so ]; then
   echo "== 32-bit-limb product (alt build)"; CKZG_B200_LIB=$PWD/c-kzg-4844_b200/build/libckzg_b200_w32.so timeout 60 python tools/pairing_probe.py 32 2>&1
 fi) | tee $OUT/pairing_probe_$TAG.log
for P in 1 0; do
  echo "== CKZG_B200_STREAM_POOL=$P"
  CKZG_B200_STREAM_POOL=$P timeout 200 python tools/e2e_dist.py 40 2>&1 | grep -v "^all"
  CKZG_B200_STREAM_POOL=$P timeout 150 python tools/e2e_probe.py 2>&1 | tail -1
  CKZG_B200_STREAM_POOL=$P PROBE_N=64 timeout 150 python tools/e2e_dist.py 40 2>&1 | grep -v "^all"
  CKZG_B200_STREAM_POOL=$P PROBE_CALLERS=2 timeout 150 python tools/e2e_probe.py 2>&1 | tail -1
done 2>&1 | tee $OUT/e2e_pool_$TAG.log
