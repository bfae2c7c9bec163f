#!/bin/bash
# N-GPU call: the multi-GPU test, the weak-scaling bench line (elbow sweep spread over the GPUs) and the sharded strong-scaling line.
N=${1:-2}; TAG=${2:-multi$N}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -m pytest tests/test_gpu_multi.py -x -q > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit $?"; cat $OUT/bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --mode sharded > $OUT/sharded_n$N.json 2> $OUT/sharded_n$N.err; echo "sharded exit $?"; cat $OUT/sharded_n$N.json
tail -3 $OUT/*.err
