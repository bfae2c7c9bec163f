#!/bin/bash
# 1 GPU: build tests (fast kernel), build micro-benchmark, ncu --set full of the fast build kernel
set -u
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab_build.py tests/test_dataframes.py -q -k "build or slab or dataframes or api_interaction" > $OUT/pytest_build.log 2>&1; echo "build pytest exit $?"; tail -12 $OUT/pytest_build.log
timeout 300 python scripts/kbench_build.py > $OUT/kbench_build.jsonl 2> $OUT/kbench_build.err; echo "kbench_build exit $?"; cat $OUT/kbench_build.jsonl; tail -2 $OUT/kbench_build.err
for sc in expression_product expression_mean; do timeout 300 python scripts/kbench_build.py --score $sc >> $OUT/kbench_build.jsonl 2>> $OUT/kbench_build.err; done; tail -2 $OUT/kbench_build.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_ccc -s 2 -c 1 -o $OUT/ncu_full_build \
    python scripts/kbench_build.py --reps 3 > $OUT/ncu_full_build.log 2>&1; echo "ncu build exit $?"
ncu -i $OUT/ncu_full_build.ncu-rep --page raw --csv > $OUT/ncu_full_build_raw.csv 2>/dev/null
ncu -i $OUT/ncu_full_build.ncu-rep --page source --csv > $OUT/ncu_full_build_source.csv 2>/dev/null
ls -la $OUT
