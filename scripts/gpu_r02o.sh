#!/bin/bash
# 1 GPU: build tests (fast kernel) + build micro-benchmark
set -u
TAG=${1:-r02o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab_build.py tests/test_dataframes.py -q -k "build or slab or dataframes or api_interaction" > $OUT/pytest_build.log 2>&1; echo "build pytest exit $?"; tail -4 $OUT/pytest_build.log
for sc in expression_gmean expression_product expression_mean; do timeout 300 python scripts/kbench_build.py --score $sc >> $OUT/kbench_build.jsonl 2>> $OUT/kbench_build.err; done; cat $OUT/kbench_build.jsonl
timeout 300 python scripts/kbench_build.py --config asd >> $OUT/kbench_build.jsonl 2>> $OUT/kbench_build.err; tail -1 $OUT/kbench_build.jsonl
