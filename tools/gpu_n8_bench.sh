#!/bin/bash
# gpurun --gpus N helper: only the N-GPU bench lines of the three BASELINE configs (no self test). usage: bash tools/gpu_n8_bench.sh <tag> <N>
tag=${1:-rX}; n=${2:-8}
out=gpurun_out/$tag
mkdir -p $out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:2}"; }
run 29521 > $out/bench_c3_n$n.json 2> $out/bench_c3_n$n.err; echo "c3 exit $?"
run 29522 --config c4 --steps 32 > $out/bench_c4_n$n.json 2> $out/bench_c4_n$n.err; echo "c4 exit $?"
run 29523 --config c5 > $out/bench_c5_n$n.json 2> $out/bench_c5_n$n.err; echo "c5 exit $?"
python - $out $n <<'PY'
import json, sys
out, n = sys.argv[1], sys.argv[2]
for c in ("c3", "c4", "c5"):
    try:
        l = json.loads(open("%s/bench_%s_n%s.json" % (out, c, n)).read().strip().splitlines()[-1])
        print(c, "Mrays/s %.1f ms %.3f e2e %.1f (%.3f ms) sha %s host_image_equal %s failed=%s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l["e2e"]["ms_per_step"],
              l["image_sha256"][:12], l["e2e"]["host_image_equal"], l.get("failed")))
    except Exception as e:
        print(c, "FAILED", e)
PY
