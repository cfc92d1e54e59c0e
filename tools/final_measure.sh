#!/bin/bash
# The round's measurement pass on one B200 (outputs under gpurun_out/): the bench line, the ncu launch list of a short run of the same
# program, and one full ncu capture each of k_scan (cfg 2), k_scan + k_xa (cfg 3's shape) and k_inflate / k_lz_resolve.
# usage: tools/final_measure.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.log; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra > gpurun_out/${tag}_plain_short.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?"
AB_MODE=0 python tools/ab_r2.py ncu1 > gpurun_out/${tag}_plain0.log 2>&1 &&
AB_MODE=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan -s 2 -c 1 -f -o gpurun_out/${tag}_k_scan python tools/ab_r2.py ncu1 > gpurun_out/${tag}_ncu0.log 2>&1; echo "k_scan capture rc=$?"
AB_MODE=1 AB_READS=30000000 python tools/ab_r2.py ncu1 > gpurun_out/${tag}_plain1.log 2>&1 &&
AB_MODE=1 AB_READS=30000000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_scan|k_xa" -s 4 -c 2 -f -o gpurun_out/${tag}_xa python tools/ab_r2.py ncu1 > gpurun_out/${tag}_ncu1.log 2>&1; echo "xa capture rc=$?"
AB_READS=50000000 python tools/ab_r2.py e2e1 > gpurun_out/${tag}_plain_e2e1.log 2>&1 &&
AB_READS=50000000 timeout 900 ncu --set full --clock-control none -k regex:"k_inflate|k_lz" -s 2 -c 2 -f -o gpurun_out/${tag}_inflate python tools/ab_r2.py e2e1 > gpurun_out/${tag}_ncu_e2e1.log 2>&1; echo "inflate capture rc=$?"
