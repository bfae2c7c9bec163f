#!/bin/bash
# per-kernel durations (ncu, cold serialised) of the masked atlas iteration: structured and random masks
set -u
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches_masked_structured.csv \
    python scripts/kbench.py --masked --structured --ranks 10 --iters 6 --reps 1 > $OUT/ncu_s.log 2>&1; echo "ncu structured exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches_masked_random.csv \
    python scripts/kbench.py --masked --ranks 10 --iters 6 --reps 1 > $OUT/ncu_r.log 2>&1; echo "ncu random exit $?"
python scripts/ncu_summarize.py launches $OUT/ncu_launches_masked_structured.csv "kbench masked structured" | tail -40
python scripts/ncu_summarize.py launches $OUT/ncu_launches_masked_random.csv "kbench masked random" | tail -40
