#!/usr/bin/env bash
# ONE ncu --set full pass over every kernel that carries a path (one call of each API, tools/prof_all.py); the raw page
# is exported and condensed by tools/ncu_summary.py, the .ncu-rep itself is dropped (hundreds of MB; gpurun_out/ only
# comes back below 64 MiB).  Usage under gpurun: bash tools/prof_all.sh <tag> [kernel-regex]
TAG=${1:-x}
K=${2:-'msm_direct_kernel|bam_fk_level1_kernel|bam_level_kernel|bam_fk_final_kernel|fk20_scalars_kernel|blob_to_cells_kernel|recover_|stage1_fused_kernel|evaluate_tree_kernel|vmsm_accumulate_kernel|vmsm_combine_kernel|vmsm_reduce_kernel|pairing_check_kernel|g1_validate_levels|vc_'}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --profile-from-start off -k "regex:$K" -c 48 -f \
    -o /tmp/prof_${TAG}_all python tools/prof_all.py > gpurun_out/ncu_${TAG}_all.log 2>&1
tail -3 gpurun_out/ncu_${TAG}_all.log | cut -c1-300
ncu -i /tmp/prof_${TAG}_all.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_all.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/raw_${TAG}_all.csv > gpurun_out/ncu_${TAG}_all_summary.txt 2>&1
wc -l gpurun_out/ncu_${TAG}_all_summary.txt
xz -9 -f gpurun_out/raw_${TAG}_all.csv
ls -la gpurun_out/ | tail -5
