#!/bin/bash
# development aid: sweep traversal tuning knobs on the bench workload
for cfg in "20 5 8" "24 5 8" "28 5 8" "16 5 8" "12 5 8" "20 3 8" "20 8 8" "20 100 8" "20 5 6" "20 5 7" "24 4 7"; do
  set -- $cfg
  TR_TUNE_FETCH=$1 TR_TUNE_POSTPONE=$2 TR_TUNE_BLOCKS=$3 timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['roofline']['class_ms_per_frame']; print('fetch $1 postpone $2 blocks $3 : %.2f ms  closest %.2f shadow %.2f shade %.2f' % (d['ms_per_step'], c['closest'], c['shadow'], c['shade']))"
done
