#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the default bench lines (ours + reference arm), a rank sweep of the streaming
# kernels, then the ncu launch list of the bench command and full captures of the streaming kernels.  Everything lands in
# gpurun_out/<tag>/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/clocks.csv 2>&1 &
SMI=$!
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json
python bench.py --impl reference --steps 2 --warmup 0 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"; cat $OUT/bench_ref.json
python scripts/kbench.py --ranks 1 2 4 5 8 10 12 13 16 20 24 25 28 30 32 --iters 50 > $OUT/kbench_ranks.jsonl 2> $OUT/kbench_ranks.err; echo "kbench exit $?"
cat $OUT/kbench_ranks.jsonl
(python scripts/kbench.py --masked --ranks 10 --iters 50; python scripts/kbench.py --masked --structured --ranks 10 --iters 50; python scripts/kbench.py --masked --structured --shape 13 1900 17 17 --ranks 10 --iters 100) > $OUT/kbench_masked.jsonl 2> $OUT/kbench_masked.err; echo "kbench masked exit $?"
kill $SMI
if [ "${NCU:-1}" = "1" ]; then
python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu --no-elbow > $OUT/bench_short.json 2> $OUT/bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --iters 20 --no-cpu --no-elbow > $OUT/ncu_bench.log 2>&1
echo "ncu launches exit $?"
python scripts/kbench.py --ranks 10 --reps 3 --quick > $OUT/kbench_short.jsonl 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'(plane|fiber)_stream' -s 2 -c 8 -o $OUT/ncu_full_stream_r10 \
    python scripts/kbench.py --ranks 10 --reps 3 --quick > $OUT/ncu_full.log 2>&1
echo "ncu full stream exit $?"
python scripts/kbench.py --ranks 20 --reps 3 --quick > $OUT/kbench_short20.jsonl 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'(plane|fiber)_mma' -s 2 -c 8 -o $OUT/ncu_full_mma_r20 \
    python scripts/kbench.py --ranks 20 --reps 3 --quick > $OUT/ncu_full20.log 2>&1
echo "ncu full mma exit $?"
fi
ls -la $OUT
