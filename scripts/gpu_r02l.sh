#!/bin/bash
# 1 GPU: mask-plan tests, masked micro-benchmark, elbow sweep with 1 / 2 / 3 host threads, default bench line
set -u
TAG=${1:-r02l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_maskplan.py tests/test_gpu_multi_runs.py -x -q > $OUT/pytest_maskplan.log 2>&1; echo "maskplan pytest exit $?"; tail -3 $OUT/pytest_maskplan.log
(timeout 300 python scripts/kbench.py --masked --ranks 4 10 20 --iters 50) > $OUT/kbench_masked.jsonl 2> $OUT/kbench_masked.err; echo "kbench masked exit $?"
cat $OUT/kbench_masked.jsonl | cut -c1-250
for T in 1 2 3; do
C2C_B200_HOST_THREADS=$T timeout 300 python scripts/elbow_breakdown.py > $OUT/elbow_threads$T.json 2> $OUT/elbow_threads$T.err; echo "elbow threads=$T exit $?"; cut -c1-300 $OUT/elbow_threads$T.json; tail -2 $OUT/elbow_threads$T.err
done
SECONDS=0
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $? after $SECONDS s"; cat $OUT/bench.json | cut -c1-300; tail -5 $OUT/bench.err
