#!/bin/bash
# The GPU calls of round 2, second session (gpurun -- bash tools/r2b_session.sh <step>): each step is what one call ran; outputs under gpurun_out/,
# the summaries that were kept under profiles/r02b_*.  Variant libraries come from tools/build_variant.sh (names in the steps).
step=$1; mkdir -p gpurun_out; t0=$(date +%s)
case "$step" in
call1)
    # round 2, second session, first GPU call: the device tests that touch the new code (k_xa out of shared memory, packed stage
    # geometry of k_scan), then the A/B of the geometries (default library and the 32-warp single-CTA variant).  Outputs: gpurun_out/r2b_*
    timeout 170 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4 \
      -k "xa_strings or adversarial or wrong_span or records_of_every_size or switches or fused_and_tuple or chunk_and_window or bench_density or matches_oracle" \
      > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b_tests.log
    timeout 130 python tools/ab_r2.py geom > gpurun_out/r2b_geom_default.log 2>&1; echo "geom default rc=$? ($(( $(date +%s) - t0 )) s)"; cat gpurun_out/r2b_geom_default.log
    ITX_LIB=$PWD/iteres_b200/csrc/variants/lib_nw32.so timeout 100 python tools/ab_r2.py geom > gpurun_out/r2b_geom_nw32.log 2>&1; echo "geom nw32 rc=$? ($(( $(date +%s) - t0 )) s)"; cat gpurun_out/r2b_geom_nw32.log
    ;;
call2)
    # second GPU call: the XA tests on the device with k_xa's register parser, cfg 3's shape timed (k_scan + k_xa), the per-kernel split
    # from an ncu duration list, and the same with the alternates switched off (variant library: what the re-derivation + staging cost)
    timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4 -k "xa_strings or adversarial or bench_density or fused_and_tuple or matches_oracle" > gpurun_out/r2b2_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b2_tests.log
    export AB_READS=50000000
    timeout 90 python tools/ab_r2.py xa > gpurun_out/r2b2_xa.log 2>&1; echo "xa rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b2_xa.log
    timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_xa|k_scan" -c 12 --csv --log-file gpurun_out/r2b2_xa_launches.csv python tools/ab_r2.py xa > gpurun_out/r2b2_xa_ncu.log 2>&1; echo "ncu rc=$? ($(( $(date +%s) - t0 )) s)"
    grep -E "k_xa|k_scan" gpurun_out/r2b2_xa_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
    ITX_LIB=$PWD/iteres_b200/csrc/variants/lib_xanoalt.so timeout 90 python tools/ab_r2.py xa > gpurun_out/r2b2_xa_noalt.log 2>&1; echo "noalt rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b2_xa_noalt.log
    ;;
call3)
    # third GPU call: XA tests with the block-reserved queue, cfg 3's shape and cfg 2's shape timed, one full ncu capture of k_scan + k_xa on the XA stream
    timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4 -k "xa_strings or adversarial or bench_density or fused_and_tuple or matches_oracle or switches" > gpurun_out/r2b3_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b3_tests.log
    export AB_READS=50000000
    timeout 90 python tools/ab_r2.py xa > gpurun_out/r2b3_xa.log 2>&1; echo "xa rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b3_xa.log
    AB_MODE=1 AB_READS=30000000 timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_scan|k_xa" -s 4 -c 2 -f -o gpurun_out/r2b3_xa python tools/ab_r2.py ncu1 > gpurun_out/r2b3_ncu.log 2>&1; echo "ncu rc=$? ($(( $(date +%s) - t0 )) s)"
    ls -la gpurun_out/r2b3_xa.ncu-rep
    ;;
call4)
    # fourth GPU call: device tests of the serial chain walk / k_xa changes, then the three record shapes with the product library and
    # with the variant that has no serial walk compiled in
    timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -n 4 -k "xa_strings or adversarial or bench_density or fused_and_tuple or matches_oracle or switches or records_of_every_size or wrong_span" > gpurun_out/r2b4_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b4_tests.log
    timeout 90 python tools/ab_r2.py three > gpurun_out/r2b4_three.log 2>&1; echo "three rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b4_three.log
    ITX_LIB=$PWD/iteres_b200/csrc/variants/lib_noserial.so timeout 90 python tools/ab_r2.py three > gpurun_out/r2b4_three_noserial.log 2>&1; echo "noserial rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b4_three_noserial.log
    ;;
call6)
    # k_overlap with a table window per warp: the tuple-path tests, then decode / overlap device time per 50 M reads
    timeout 75 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -n 6 -k "fused_and_tuple or matches_oracle or xa_strings or adversarial or wrong_span or records_of_every_size or rmdup or nested or overlap_kernel or filter_mode or kat4 or kat2 or kat3" > gpurun_out/r2b6_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/r2b6_tests.log
    AB_TUPLE_ONE=1 AB_READS=50000000 timeout 40 python tools/ab_r2.py tuple > gpurun_out/r2b6_tuple.log 2>&1; echo "tuple rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v Warning gpurun_out/r2b6_tuple.log
    ;;
final)
    # the session's last GPU call: the whole device suite, smoke(), the bench line of the final code
    timeout 120 python -m pytest tests -m gpu -q -n 6 > gpurun_out/r2b_final_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/r2b_final_tests.log
    timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_final_smoke.log 2>&1; echo "smoke rc=$? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/r2b_final_smoke.log
    timeout 180 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_final_bench.json 2> gpurun_out/r2b_final_bench.log; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; tail -c 600 gpurun_out/r2b_final_bench.json; tail -5 gpurun_out/r2b_final_bench.log
    ;;
*) echo "usage: $0 call1|call2|call3|call4|call6|final" ;;
esac
