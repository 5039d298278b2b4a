#!/bin/bash
# development aid: sweep traversal tuning knobs on the bench workload
for cfg in "20 5 8" "20 2 8" "20 1 8" "24 2 8" "16 2 8"; do
  set -- $cfg
  TR_TUNE_FETCH=$1 TR_TUNE_POSTPONE=$2 TR_TUNE_BLOCKS=$3 timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['roofline']['class_ms_per_frame']; print('fetch $1 postpone $2 blocks $3 : %.2f ms  closest %.2f shadow %.2f shade %.2f nodes %.1f' % (d['ms_per_step'], c['closest'], c['shadow'], c['shade'], d['roofline']['nodes_per_ray']['closest']))"
done
