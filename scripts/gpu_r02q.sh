#!/bin/bash
set -u
TAG=${1:-r02q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/bin_probe.py --ranks 25 3 > $OUT/bin_probe.jsonl 2> $OUT/bin_probe.err; echo "probe exit $?"; cat $OUT/bin_probe.jsonl; tail -2 $OUT/bin_probe.err
timeout 300 python scripts/bin_probe.py --ranks 14 14 >> $OUT/bin_probe.jsonl 2>> $OUT/bin_probe.err; tail -1 $OUT/bin_probe.jsonl
timeout 300 python scripts/bin_probe.py --ranks 7 7 7 7 >> $OUT/bin_probe.jsonl 2>> $OUT/bin_probe.err; tail -1 $OUT/bin_probe.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/ncu_launches_bin.csv \
    python scripts/bin_probe.py --ranks 25 3 --iters 10 --reps 1 > $OUT/ncu_bin.log 2>&1; echo "ncu exit $?"
python scripts/ncu_summarize.py launches $OUT/ncu_launches_bin.csv "bin_probe --ranks 25 3 --iters 10" | grep -v "at::\|cutlass\|internal::" | tail -30
timeout 600 python scripts/kbench_build.py > $OUT/kbench_build.jsonl 2>&1; cat $OUT/kbench_build.jsonl
