#!/bin/bash
# gpurun helper: GPU parity suite + the bench line. usage: bash tools/gpu_tests_and_bench.sh <tag> [pytest args]
tag=${1:-rX}; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q --durations=15 "$@" > $out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/pytest_gpu.log
tail -25 $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
echo "bench exit $?"
tail -c 1500 $out/bench_n1.json; tail -5 $out/bench_n1.err
