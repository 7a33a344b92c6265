#!/usr/bin/env bash
# Short GPU-box visit for the EIP-7594 side: the cells/recover/verify-cell parity tests, then their bench extras.
set -u
TAG=${1:-c}
OUT=gpurun_out
mkdir -p $OUT
CKZG_B200_DEBUG=1 timeout 900 python -m pytest tests -q -m gpu -x -k "cells or recover or golden_vectors" 2>&1 | tail -5 | tee $OUT/pytest_$TAG.log
timeout 900 python bench.py --steps 3 --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_$TAG.log
python - <<PY
import json
d=json.loads([x for x in open("$OUT/bench_$TAG.log") if x.startswith("{")][-1])
for k,v in d["extra"].items():
    if "cell" in k or "recover" in k: print(k, v)
PY
