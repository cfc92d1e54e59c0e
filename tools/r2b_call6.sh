#!/bin/bash
# k_overlap with a table window per warp: the tuple-path tests, then decode / overlap device time per 50 M reads
mkdir -p gpurun_out
t0=$(date +%s)
timeout 75 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -n 6 -k "fused_and_tuple or matches_oracle or xa_strings or adversarial or wrong_span or records_of_every_size or rmdup or nested or overlap_kernel or filter_mode or kat4 or kat2 or kat3" > gpurun_out/r2b6_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b6_tests.log
AB_TUPLE_ONE=1 AB_READS=50000000 timeout 40 python tools/ab_r2.py tuple > gpurun_out/r2b6_tuple.log 2>&1; echo "tuple rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b6_tuple.log
