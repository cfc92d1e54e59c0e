#!/bin/bash
# second GPU call: the XA tests on the device with k_xa's register parser, cfg 3's shape timed (k_scan + k_xa), the per-kernel split
# from an ncu duration list, and the same with the alternates switched off (variant library: what the re-derivation + staging cost)
mkdir -p gpurun_out
t0=$(date +%s)
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4 -k "xa_strings or adversarial or bench_density or fused_and_tuple or matches_oracle" > gpurun_out/r2b2_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b2_tests.log
export AB_READS=50000000
timeout 90 python tools/ab_r2.py xa > gpurun_out/r2b2_xa.log 2>&1; echo "xa rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b2_xa.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_xa|k_scan" -c 12 --csv --log-file gpurun_out/r2b2_xa_launches.csv python tools/ab_r2.py xa > gpurun_out/r2b2_xa_ncu.log 2>&1; echo "ncu rc=$? ($(( $(date +%s) - t0 )) s)"
grep -E "k_xa|k_scan" gpurun_out/r2b2_xa_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
ITX_LIB=$PWD/iteres_b200/csrc/variants/lib_xanoalt.so timeout 90 python tools/ab_r2.py xa > gpurun_out/r2b2_xa_noalt.log 2>&1; echo "noalt rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b2_xa_noalt.log
