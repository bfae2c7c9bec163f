#!/bin/bash
# Investigation call: memory / issue-rate ceilings, e2e breakdown, per-kernel launch list at high ranks, full ncu captures
# (with source) of the small update kernels.
set -u
TAG=${1:-probe}
OUT=gpurun_out/$TAG
mkdir -p $OUT
build/membench > $OUT/membench.jsonl 2>&1; echo "membench exit $?"
build/mmabench > $OUT/mmabench.jsonl 2>&1; echo "mmabench exit $?"; cat $OUT/mmabench.jsonl
python scripts/e2e_breakdown.py > $OUT/e2e_breakdown.json 2> $OUT/e2e_breakdown.err; echo "e2e exit $?"; cat $OUT/e2e_breakdown.json
python scripts/kbench.py --ranks 24 25 --reps 3 --iters 3 > $OUT/kbench_hi.jsonl 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/ncu_launches_hi.csv \
    python scripts/kbench.py --ranks 24 25 --reps 3 --iters 3 > $OUT/ncu_hi.log 2>&1
echo "ncu hi exit $?"
python scripts/kbench.py --ranks 10 --reps 3 --iters 3 > $OUT/kbench_short.jsonl 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'(lead_update|trail_update|fiber_reduce2|fiber_stream)' -s 10 -c 5 -o $OUT/ncu_full_small_r10 \
    python scripts/kbench.py --ranks 10 --reps 3 --iters 3 > $OUT/ncu_full_small.log 2>&1
echo "ncu small exit $?"
ls -la $OUT
