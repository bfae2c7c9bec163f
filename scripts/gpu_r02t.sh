#!/bin/bash
set -u
TAG=${1:-r02t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ncp_golden.py -q -k "recon_sums or config4 or full_size or mma or api_ or reference_driver" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest.log
timeout 300 python scripts/sums_bench.py --ranks 10 13 16 20 25 32 > $OUT/sums_bench.jsonl 2> $OUT/sums_bench.err; echo "sums_bench exit $?"; cat $OUT/sums_bench.jsonl | cut -c1-300; tail -2 $OUT/sums_bench.err
timeout 300 python scripts/bin_probe.py --ranks 25 3 > $OUT/bin_probe.jsonl 2> $OUT/bin_probe.err; cat $OUT/bin_probe.jsonl
timeout 300 python scripts/elbow_breakdown.py > $OUT/elbow.json 2> $OUT/elbow.err; echo "elbow exit $?"; cut -c1-200 $OUT/elbow.json
