#!/bin/bash
# Coupled factorization: GPU parity tests + micro-benchmark.   gpurun --timeout 900 -- 'bash scripts/gpu_coupled.sh [tag]'
set -u
TAG=${1:-coupled}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_coupled.py -x -q > $OUT/pytest_coupled.log 2>&1; echo "pytest exit $?"
tail -40 $OUT/pytest_coupled.log
timeout 300 python scripts/coupled_bench.py > $OUT/coupled_bench.jsonl 2> $OUT/coupled_bench.err; echo "bench exit $?"
cat $OUT/coupled_bench.jsonl; tail -5 $OUT/coupled_bench.err
