#!/bin/bash
# N-GPU call: the multi-GPU tests, the default bench line under torchrun (weak scaling + the sharded / config-5 sub-records).
N=${1:-2}; TAG=${2:-multi$N}; EXTRA=${3:-}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded_fused.py $EXTRA -q > $OUT/pytest_multi.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -4 $OUT/pytest_multi.log
SECONDS=0
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit $? after $SECONDS s"; cat $OUT/bench_n$N.json
grep -v "^\s*$" $OUT/bench_n$N.err | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -8
