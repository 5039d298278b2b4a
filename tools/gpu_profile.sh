#!/bin/bash
# gpurun helper: the ncu captures profiles/ is built from (numbers printed by a run under ncu are never bench values).
# usage: bash tools/gpu_profile.sh <tag> [bench args]
tag=${1:-rX}; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_raw.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $out/ncu_launch_bench.log 2>&1
# (frame by frame, so that the 13 launches after the skip are ONE steady-state frame; ncu serialises the launches anyway)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf_ --launch-skip 39 -c 13 -f -o $out/full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --frames-in-flight 1 "$@" > $out/ncu_full_bench.log 2>&1
python tools/summarize_profile.py launches $out/launches_raw.csv $out/launch_shares.csv
python tools/summarize_profile.py full $out/full.ncu-rep $out/ncu_full_top_kernels.csv
python tools/ncu_table.py $out/full.ncu-rep > $out/ncu_pipes_and_stalls.txt 2>&1
cat $out/launch_shares.csv | head -30
cat $out/ncu_pipes_and_stalls.txt | cut -c1-260
ls -la $out
