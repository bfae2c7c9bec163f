#!/bin/bash
set -u
TAG=${1:-r02y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "mma_passes or config4 or mma_run" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest.log
timeout 200 python scripts/kbench.py --ranks 17 20 24 --iters 30 2>/dev/null | cut -c1-330
C2C_B200_NO_MMA_N24=1 timeout 200 python scripts/kbench.py --ranks 20 --iters 30 2>/dev/null | cut -c1-330
