#!/bin/bash
# 1 GPU: full GPU suite (no -x), masked micro-benchmarks, default bench line, launch list of the masked random iteration
set -u
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SECONDS=0
timeout 1700 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -15 $OUT/pytest_gpu.log
(timeout 300 python scripts/kbench.py --masked --ranks 4 10 20 --iters 50; timeout 300 python scripts/kbench.py --masked --structured --ranks 10 --iters 50) > $OUT/kbench_masked.jsonl 2> $OUT/kbench_masked.err; echo "kbench masked exit $?"
cat $OUT/kbench_masked.jsonl; tail -3 $OUT/kbench_masked.err
SECONDS=0
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $? after $SECONDS s"; cat $OUT/bench.json | cut -c1-600; tail -5 $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches_masked_random.csv \
    python scripts/kbench.py --masked --ranks 10 --iters 6 --reps 1 > $OUT/ncu_r.log 2>&1; echo "ncu random exit $?"
python scripts/ncu_summarize.py launches $OUT/ncu_launches_masked_random.csv "kbench masked random" | grep -v "at::\|cutlass\|internal::" | tail -30
