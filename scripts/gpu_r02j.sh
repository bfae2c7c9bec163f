#!/bin/bash
# 1 GPU: ncu --set full (with source) of the residual-list correction kernels on the masked random-20% atlas iteration
set -u
TAG=${1:-r02j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/kbench.py --masked --ranks 10 --iters 4 --reps 1 > $OUT/kbench_short.jsonl 2>&1; echo "kbench exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_list -s 4 -c 4 -o $OUT/ncu_full_corr_list \
    python scripts/kbench.py --masked --ranks 10 --iters 4 --reps 1 > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i $OUT/ncu_full_corr_list.ncu-rep --page raw --csv > $OUT/ncu_full_corr_list_raw.csv 2>/dev/null
ncu -i $OUT/ncu_full_corr_list.ncu-rep --page source --csv > $OUT/ncu_full_corr_list_source.csv 2>/dev/null
ls -la $OUT
