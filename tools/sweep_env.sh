#!/bin/bash
# gpurun helper: A/B the bench under different TR_TUNE_* environments. usage: bash tools/sweep_env.sh <tag> "ENV1=.. ENV2=.." "ENV.." ...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --no-cpu-baseline --steps 10 > $out/bench_$i.json 2> $out/bench_$i.err
  python - "$cfg" $out/bench_$i.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = l["roofline"]
    print("%-70s ms %.3f  closest %.2f shadow %.2f shade %.2f other %.2f  nodes %.2f/%.2f tris %.2f/%.2f sha %s" % (
        sys.argv[1], l["ms_per_step"], r["class_ms_per_frame"]["closest"], r["class_ms_per_frame"]["shadow"], r["class_ms_per_frame"]["shade"],
        r["class_ms_per_frame"]["other"], r["nodes_per_ray"]["closest"], r["nodes_per_ray"]["shadow"], r["tris_per_ray"]["closest"], r["tris_per_ray"]["shadow"],
        l["image_sha256"][:12]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
