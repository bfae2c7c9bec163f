#!/bin/bash
# round-2 state check: full GPU suite, small-tensor and masked micro-benchmarks, default bench line, masked launch lists
set -u
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
(timeout 300 python scripts/kbench_small.py --workload balf; timeout 300 python scripts/kbench_small.py --workload toy; timeout 300 python scripts/kbench_small.py --workload asd) > $OUT/kbench_small.jsonl 2> $OUT/kbench_small.err; echo "kbench small exit $?"
cat $OUT/kbench_small.jsonl; tail -5 $OUT/kbench_small.err
(timeout 300 python scripts/kbench.py --masked --structured --ranks 10 20 --iters 50; timeout 300 python scripts/kbench.py --masked --ranks 4 10 20 --iters 50; timeout 300 python scripts/kbench.py --masked --structured --shape 13 1900 17 17 --ranks 10 --iters 100; timeout 300 python scripts/kbench.py --ranks 10 20 --iters 50) > $OUT/kbench_masked.jsonl 2> $OUT/kbench_masked.err; echo "kbench masked exit $?"
cat $OUT/kbench_masked.jsonl; tail -5 $OUT/kbench_masked.err
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches_masked_structured.csv \
    python scripts/kbench.py --masked --structured --ranks 10 --iters 6 --reps 1 > $OUT/ncu_s.log 2>&1; echo "ncu structured exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches_masked_random.csv \
    python scripts/kbench.py --masked --ranks 10 --iters 6 --reps 1 > $OUT/ncu_r.log 2>&1; echo "ncu random exit $?"
python scripts/ncu_summarize.py launches $OUT/ncu_launches_masked_structured.csv "kbench masked structured" | grep -v "at::\|cutlass\|internal::" | tail -30
python scripts/ncu_summarize.py launches $OUT/ncu_launches_masked_random.csv "kbench masked random" | grep -v "at::\|cutlass\|internal::" | tail -30
