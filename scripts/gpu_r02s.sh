#!/bin/bash
# 1 GPU: full GPU suite, smoke, default bench line, reference arm (short), ncu launch list of the bench command
set -u
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -6 $OUT/pytest_gpu.log
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
SECONDS=0
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $? after $SECONDS s"; cut -c1-700 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu --no-elbow --no-extras --no-sharded --no-coupled > $OUT/bench_short.json 2> $OUT/bench_short.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu --no-elbow --no-extras --no-sharded --no-coupled > $OUT/ncu_bench.log 2>&1
echo "ncu launches exit $?"
python scripts/ncu_summarize.py launches $OUT/ncu_launches_bench.csv "bench short" | grep -v "at::\|cutlass\|internal::" | tail -16
