#!/bin/bash
# development aid (GPU box): sweep environment tuning knobs on the bench workload
# usage: tools/sweep.sh "VAR=val VAR2=val" "VAR=val" ...
for cfg in "$@"; do
  env $cfg timeout 200 python bench.py --no-cpu-baseline --steps ${STEPS:-6} --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['roofline']['class_ms_per_frame']; r=d['roofline']; print('%-44s %.2f ms  closest %.2f shadow %.2f shade %.2f | nodes %.1f/%.1f tris %.2f/%.2f build %.1f ms' % ('$cfg', d['ms_per_step'], c['closest'], c['shadow'], c['shade'], r['nodes_per_ray']['closest'], r['nodes_per_ray']['shadow'], r['tris_per_ray']['closest'], r['tris_per_ray']['shadow'], d['config']['bvh_build_ms']))"
done
