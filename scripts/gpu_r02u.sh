#!/bin/bash
# 1 GPU: new maskplan test + reference arm of the bench (bounded) -- quick checks
set -u
TAG=${1:-r02u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_maskplan.py -q -k "several_chunks or stop_rule or all_ones" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
timeout 300 python scripts/kbench.py --masked --ranks 10 --iters 50 > $OUT/kbench_masked.jsonl 2>&1; cut -c1-250 $OUT/kbench_masked.jsonl
