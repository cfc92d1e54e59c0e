#!/bin/bash
# compile-time variants of libiteres_gpu.so (iteres_b200/csrc/variants/lib_<name>.so) through the bench's device-resident value
# usage: tools/ab_variants.sh <tag> name1 name2 ...   ("default" = the library in the tree)
tag=$1; shift
mkdir -p gpurun_out
for name in "$@"; do
    lib=""; [ "$name" != default ] && lib="$PWD/iteres_b200/csrc/variants/lib_$name.so"
    ITX_LIB=$lib python bench.py --no-e2e --no-cpu-baseline --no-extra --steps 20 --warmup 3 $EXTRA > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.log
    python - "$name" gpurun_out/${tag}_$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-14s ms/step %.3f  k_scan %.3f ms  frac %.3f  frac(stream) %.3f" % (sys.argv[1], d["ms_per_step"], r["kernel_ms_per_step"], r["frac"], r["frac_stream_bytes_only"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
