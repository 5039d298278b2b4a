#!/bin/bash
# One gpurun call that produces everything profiles/ is built from: GPU parity suite, bench line,
# reference-arm line, ncu launch list, ncu --set full of one steady-state frame.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-rX}
out=gpurun_out/$tag
mkdir -p $out
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $out/pytest_gpu.log
  tail -3 $out/pytest_gpu.log
fi
timeout 300 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
tail -c 600 $out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2> $out/bench_ref.err
tail -c 300 $out/bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_raw.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_ --launch-skip 13 -c 13 -f -o $out/full \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_full_bench.log 2>&1
ls -la $out
