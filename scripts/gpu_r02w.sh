#!/bin/bash
set -u
TAG=${1:-r02w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SECONDS=0
timeout 600 python bench.py --no-elbow --no-extras --no-sharded --no-coupled --no-cpu > $OUT/bench_short.json 2> $OUT/bench_short.err; echo "bench exit $? after $SECONDS s"; tail -3 $OUT/bench_short.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02w/bench_short.json").read().strip().splitlines()[-1])
print(json.dumps(d["roofline"])[:1500]); print(json.dumps(d["rank20"])[:900]); print(d["value"], d["e2e"]["value"])
PY
