#!/bin/bash
# Coupled factorization: GPU parity tests, micro-benchmark (two-stream and one-stream iteration), ncu launch list.
#   gpurun --timeout 900 -- 'bash scripts/gpu_coupled.sh [tag]'
set -u
TAG=${1:-coupled}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_coupled.py -x -q > $OUT/pytest_coupled.log 2>&1; echo "pytest exit $?"
tail -40 $OUT/pytest_coupled.log
timeout 300 python scripts/coupled_bench.py > $OUT/coupled_bench.jsonl 2> $OUT/coupled_bench.err; echo "bench exit $?"
cat $OUT/coupled_bench.jsonl; tail -5 $OUT/coupled_bench.err
C2C_B200_COUPLED_ONE_STREAM=1 timeout 300 python scripts/coupled_bench.py > $OUT/coupled_bench_one_stream.jsonl 2>> $OUT/coupled_bench.err; echo "bench(1 stream) exit $?"
cat $OUT/coupled_bench_one_stream.jsonl
if [ "${NCU:-1}" = "1" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches_coupled.csv \
    python scripts/coupled_bench.py --ranks 10 --iters 5 10 > $OUT/ncu_coupled.log 2>&1
echo "ncu launches exit $?"
fi
