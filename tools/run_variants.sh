#!/bin/bash
# development aid (GPU box): bench every build/variants/*.so, print ms per kernel class
for so in build/variants/*.so; do
  TAURAY_B200_LIB=$PWD/$so timeout 200 python bench.py --no-cpu-baseline --steps ${STEPS:-6} --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['roofline']['class_ms_per_frame']; print('%-28s %.2f ms  closest %.2f shadow %.2f shade %.2f other %.2f' % ('$(basename $so .so)', d['ms_per_step'], c['closest'], c['shadow'], c['shade'], c['other']))"
done
