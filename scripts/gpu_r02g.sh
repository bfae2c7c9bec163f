#!/bin/bash
# 1 GPU: full GPU suite, default bench line (timed), ncu --set full of the residual-list correction kernels (masked, random 20 %)
set -u
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1700 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
SECONDS=0
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $? after $SECONDS s"; cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_list -c 4 -o $OUT/ncu_full_corr_list \
    python scripts/kbench.py --masked --ranks 10 --iters 4 --reps 1 > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i $OUT/ncu_full_corr_list.ncu-rep --page raw --csv > $OUT/ncu_full_corr_list_raw.csv 2>/dev/null
ls -la $OUT
