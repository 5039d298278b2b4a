#!/bin/bash
# gpurun --gpus N helper: the multi-GPU test worker + the N-GPU bench line. usage: bash tools/gpu_multi.sh <tag> <N>
tag=${1:-rX}; n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > $out/pytest_multi.log 2>&1
echo "pytest exit $?" >> $out/pytest_multi.log
tail -30 $out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > $out/bench_n$n.json 2> $out/bench_n$n.err
echo "bench exit $?"
tail -c 2500 $out/bench_n$n.json; tail -5 $out/bench_n$n.err
