#!/bin/bash
# e2e (BGZF file -> tables) with different inflate launch-group sizes.  usage: tools/ab_e2e.sh <tag>
tag=${1:-e2e}
for g in 8192 16384 4096; do
    ITX_INF_GROUP=$g python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 3 > gpurun_out/${tag}_g$g.json 2> gpurun_out/${tag}_g$g.log
    python - $g gpurun_out/${tag}_g$g.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1]); e = d["e2e"]
    print("group %s: e2e %.1f M reads/s, %.1f ms/step, inflate %.1f ms, scan stream %.1f ms" % (sys.argv[1], e["value"] / 1e6, 1e3 * e["s_per_step"], e["inflate_ms"], e["scan_stream_ms"]))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
