#!/bin/bash
# round 2, second session, first GPU call: the device tests that touch the new code (k_xa out of shared memory, packed stage
# geometry of k_scan), then the A/B of the geometries (default library and the 32-warp single-CTA variant).  Outputs: gpurun_out/r2b_*
mkdir -p gpurun_out
t0=$(date +%s)
timeout 170 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4 \
  -k "xa_strings or adversarial or wrong_span or records_of_every_size or switches or fused_and_tuple or chunk_and_window or bench_density or matches_oracle" \
  > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b_tests.log
timeout 130 python tools/ab_r2.py geom > gpurun_out/r2b_geom_default.log 2>&1; echo "geom default rc=$? ($(( $(date +%s) - t0 )) s)"; cat gpurun_out/r2b_geom_default.log
ITX_LIB=$PWD/iteres_b200/csrc/variants/lib_nw32.so timeout 100 python tools/ab_r2.py geom > gpurun_out/r2b_geom_nw32.log 2>&1; echo "geom nw32 rc=$? ($(( $(date +%s) - t0 )) s)"; cat gpurun_out/r2b_geom_nw32.log
