#!/bin/bash
# masked-path check: mask-plan tests + masked parity tests, then the masked micro-benchmarks
set -u
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_maskplan.py -x -q > $OUT/pytest_maskplan.log 2>&1; echo "pytest maskplan exit $?"; tail -15 $OUT/pytest_maskplan.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "masked or odd or config3" > $OUT/pytest_masked.log 2>&1; echo "pytest masked exit $?"; tail -5 $OUT/pytest_masked.log
(timeout 300 python scripts/kbench.py --masked --structured --ranks 10 20 --iters 50; timeout 300 python scripts/kbench.py --masked --ranks 4 10 20 --iters 50; timeout 300 python scripts/kbench.py --masked --structured --shape 13 1900 17 17 --ranks 10 --iters 100; timeout 300 python scripts/kbench.py --ranks 10 --iters 50) > $OUT/kbench_masked.jsonl 2> $OUT/kbench_masked.err; echo "kbench masked exit $?"
cat $OUT/kbench_masked.jsonl; tail -5 $OUT/kbench_masked.err
