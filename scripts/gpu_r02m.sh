#!/bin/bash
# 1 GPU: tensor-build micro-benchmark + ncu --set full of build_ccc; ncu --set full of the tensor-pipe passes at rank 20
set -u
TAG=${1:-r02m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/kbench_build.py > $OUT/kbench_build.jsonl 2> $OUT/kbench_build.err; echo "kbench_build exit $?"; cat $OUT/kbench_build.jsonl; tail -2 $OUT/kbench_build.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_ccc -s 2 -c 2 -o $OUT/ncu_full_build \
    python scripts/kbench_build.py --reps 3 > $OUT/ncu_full_build.log 2>&1; echo "ncu build exit $?"
ncu -i $OUT/ncu_full_build.ncu-rep --page raw --csv > $OUT/ncu_full_build_raw.csv 2>/dev/null
ncu -i $OUT/ncu_full_build.ncu-rep --page source --csv > $OUT/ncu_full_build_source.csv 2>/dev/null
timeout 300 python scripts/kbench.py --ranks 20 --reps 3 --quick > $OUT/kbench_short20.jsonl 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'(plane|fiber)_mma' -s 2 -c 4 -o $OUT/ncu_full_mma_r20 \
    python scripts/kbench.py --ranks 20 --reps 3 --quick > $OUT/ncu_full20.log 2>&1
echo "ncu full mma exit $?"
ncu -i $OUT/ncu_full_mma_r20.ncu-rep --page raw --csv > $OUT/ncu_full_mma_r20_raw.csv 2>/dev/null
ls -la $OUT
