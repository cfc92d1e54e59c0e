#!/bin/bash
# fourth GPU call: device tests of the serial chain walk / k_xa changes, then the three record shapes with the product library and
# with the variant that has no serial walk compiled in
mkdir -p gpurun_out
t0=$(date +%s)
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4 -k "xa_strings or adversarial or bench_density or fused_and_tuple or matches_oracle or switches or records_of_every_size or wrong_span" > gpurun_out/r2b4_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b4_tests.log
timeout 90 python tools/ab_r2.py three > gpurun_out/r2b4_three.log 2>&1; echo "three rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b4_three.log
ITX_LIB=$PWD/iteres_b200/csrc/variants/lib_noserial.so timeout 90 python tools/ab_r2.py three > gpurun_out/r2b4_three_noserial.log 2>&1; echo "noserial rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b4_three_noserial.log
