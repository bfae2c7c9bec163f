#!/bin/bash
set -u
TAG=${1:-r02v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for ep in 8 16 24 32 48; do
  echo "epoch $ep"; C2C_B200_MMA_EPOCH=$ep timeout 200 python scripts/kbench.py --ranks 20 28 --iters 30 2>/dev/null | cut -c1-330 | tee -a $OUT/epoch_$ep.jsonl
done
