#!/bin/bash
set -u
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_multi_runs.py -x -q > $OUT/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -15 $OUT/pytest_multi.log
timeout 900 python -m pytest tests/test_gpu_maskplan.py -x -q > $OUT/pytest_maskplan.log 2>&1; echo "pytest maskplan exit $?"; tail -5 $OUT/pytest_maskplan.log
(timeout 300 python scripts/kbench_small.py --workload balf; timeout 300 python scripts/kbench_small.py --workload toy; timeout 300 python scripts/kbench_small.py --workload asd) > $OUT/kbench_small.jsonl 2> $OUT/kbench_small.err; echo "kbench small exit $?"
cat $OUT/kbench_small.jsonl; tail -5 $OUT/kbench_small.err
(timeout 300 python scripts/kbench.py --masked --structured --ranks 10 20 --iters 50; timeout 300 python scripts/kbench.py --masked --ranks 10 --iters 50; timeout 300 python scripts/kbench.py --masked --structured --shape 13 1900 17 17 --ranks 10 --iters 100) > $OUT/kbench_masked.jsonl 2> $OUT/kbench_masked.err; echo "kbench masked exit $?"
cat $OUT/kbench_masked.jsonl; tail -5 $OUT/kbench_masked.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches_masked_structured.csv \
    python scripts/kbench.py --masked --structured --ranks 10 --iters 6 --reps 1 > $OUT/ncu_s.log 2>&1; echo "ncu structured exit $?"
python scripts/ncu_summarize.py launches $OUT/ncu_launches_masked_structured.csv "kbench masked structured" | grep -v "at::\|cutlass\|internal::" | tail -30
