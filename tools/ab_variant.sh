python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu37.log 2>&1; tail -3 gpurun_out/pytest_gpu37.log
tools/ab_scan.sh r01l
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01l_w14.json').read().strip().splitlines()[-1])
print("tuple path", d["roofline"]["all_kernels"]["tuple_path"])
PY
cp iteres_b200/csrc/libiteres_gpu.so /tmp/keep.so; cp iteres_b200/csrc/variant_w16.so iteres_b200/csrc/libiteres_gpu.so
python -m pytest tests -m gpu -x -q -k "golden or synth or fused" > gpurun_out/pytest_gpu37_w16.log 2>&1; tail -2 gpurun_out/pytest_gpu37_w16.log
python bench.py --no-e2e --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/r01l_w16.json 2> gpurun_out/r01l_w16.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01l_w16.json').read().strip().splitlines()[-1])
print("w16 variant: k_scan %.3f ms frac %.3f" % (d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"]))
PY
cp /tmp/keep.so iteres_b200/csrc/libiteres_gpu.so
