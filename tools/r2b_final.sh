#!/bin/bash
# the session's last GPU call: the whole device suite, smoke(), the bench line of the final code
mkdir -p gpurun_out
t0=$(date +%s)
timeout 120 python -m pytest tests -m gpu -q -n 6 > gpurun_out/r2b_final_tests.log 2>&1; echo "tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/r2b_final_tests.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_final_smoke.log 2>&1; echo "smoke rc=$? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/r2b_final_smoke.log
timeout 180 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_final_bench.json 2> gpurun_out/r2b_final_bench.log; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; tail -c 600 gpurun_out/r2b_final_bench.json; tail -5 gpurun_out/r2b_final_bench.log
