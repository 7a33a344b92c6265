#!/usr/bin/env bash
# compute-sanitizer over one small pass of every API entry (tools/sanitize.py): memcheck, racecheck, synccheck.
# Small fixed-base tables (the sanitizer slows kernels 10-100x).  Usage under gpurun: bash tools/gpu_sanitize.sh [tag]
TAG=${1:-s}
OUT=gpurun_out
mkdir -p $OUT
export CKZG_B200_COMMIT_WINDOW=10 CKZG_B200_FK_WINDOW=8 CKZG_B200_TAIL_PIECES=8
for tool in ${SANITIZE_TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool" | tee -a $OUT/sanitize_$TAG.log
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize pass ok|hazard|Error|error" | head -20 | tee -a $OUT/sanitize_$TAG.log
done
