#!/bin/bash
# third GPU call: XA tests with the block-reserved queue, cfg 3's shape and cfg 2's shape timed, one full ncu capture of k_scan + k_xa on the XA stream
mkdir -p gpurun_out
t0=$(date +%s)
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4 -k "xa_strings or adversarial or bench_density or fused_and_tuple or matches_oracle or switches" > gpurun_out/r2b3_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b3_tests.log
export AB_READS=50000000
timeout 90 python tools/ab_r2.py xa > gpurun_out/r2b3_xa.log 2>&1; echo "xa rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b3_xa.log
AB_MODE=1 AB_READS=30000000 timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_scan|k_xa" -s 4 -c 2 -f -o gpurun_out/r2b3_xa python tools/ab_r2.py ncu1 > gpurun_out/r2b3_ncu.log 2>&1; echo "ncu rc=$? ($(( $(date +%s) - t0 )) s)"
ls -la gpurun_out/r2b3_xa.ncu-rep
