#!/bin/bash
# 1 GPU: mask-plan tests first (new residual-list kernels), then the full GPU suite, masked micro-benchmarks, launch list
set -u
TAG=${1:-r02k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_maskplan.py -x -q > $OUT/pytest_maskplan.log 2>&1; rc=$?; echo "maskplan pytest exit $rc after $SECONDS s"; tail -15 $OUT/pytest_maskplan.log
(timeout 300 python scripts/kbench.py --masked --ranks 4 10 20 --iters 50; timeout 300 python scripts/kbench.py --masked --structured --ranks 10 --iters 50) > $OUT/kbench_masked.jsonl 2> $OUT/kbench_masked.err; echo "kbench masked exit $?"
cat $OUT/kbench_masked.jsonl; tail -3 $OUT/kbench_masked.err
if [ $rc -ne 0 ]; then exit 1; fi
SECONDS=0
timeout 1700 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -8 $OUT/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches_masked_random.csv \
    python scripts/kbench.py --masked --ranks 10 --iters 6 --reps 1 > $OUT/ncu_r.log 2>&1; echo "ncu random exit $?"
python scripts/ncu_summarize.py launches $OUT/ncu_launches_masked_random.csv "kbench masked random" | grep -v "at::\|cutlass\|internal::" | tail -30
