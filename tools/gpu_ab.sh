#!/bin/bash
# gpurun helper: A/B every build/variants/*.so under a list of TR_TUNE_* environments: bench line (ms per kernel class,
# image hash) and the 1/8-strip timing. usage: bash tools/gpu_ab.sh <tag> "ENV=.." "ENV=.." ...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
for so in build/variants/*.so; do
  name=$(basename $so .so)
  for cfg in "$@"; do
    key="$name $cfg"
    env $cfg TAURAY_B200_LIB=$PWD/$so timeout 300 python bench.py --no-cpu-baseline --steps ${STEPS:-10} > "$out/bench_${name}_${cfg//[^A-Za-z0-9=]/_}.json" 2> $out/err.log
    python - "$key" "$out/bench_${name}_${cfg//[^A-Za-z0-9=]/_}.json" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = l["roofline"]
    print("%-40s ms %.3f  closest %.2f shadow %.2f shade %.2f other %.2f  nodes %.2f/%.2f tris %.2f/%.2f sha %s" % (
        sys.argv[1], l["ms_per_step"], r["class_ms_per_frame"]["closest"], r["class_ms_per_frame"]["shadow"], r["class_ms_per_frame"]["shade"],
        r["class_ms_per_frame"]["other"], r["nodes_per_ray"]["closest"], r["nodes_per_ray"]["shadow"], r["tris_per_ray"]["closest"], r["tris_per_ray"]["shadow"],
        l["image_sha256"][:12]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
    if [ -n "$STRIP" ]; then
      for w in $STRIP; do echo -n "   strip: "; env $cfg TAURAY_B200_LIB=$PWD/$so timeout 200 python tools/strip_bench.py $w 20 2>&1 | tail -1; done
    fi
  done
done 2>&1 | tee $out/summary.txt
