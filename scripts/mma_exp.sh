#!/bin/bash
# timing experiments on the tensor-core passes: which role of the pipeline limits them
OUT=gpurun_out/${1:-mmaexp}; mkdir -p $OUT
run() { echo "== $*"; env "$@" timeout 120 python scripts/kbench.py --ranks ${RANKS:-10} --reps 10 --quick --tag "$*" 2>&1 | tail -1; }
{
for d in 0 1 2 3 8; do run C2C_B200_MMA_DBG=$d; done     # (masks with bit 4 -- no MMAs issued -- are not run: the epilogue would wait forever)
run C2C_B200_MMA_DBG=0 C2C_B200_MMA_EPOCH=64
} > $OUT/exp.log 2>&1
cat $OUT/exp.log
