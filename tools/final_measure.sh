#!/bin/bash
# The round's measurement pass on one B200: the bench line, the launch list of the same command, one full ncu capture of
# k_scan at the bench's size.  usage: tools/final_measure.sh <tag>   (outputs under gpurun_out/)
tag=${1:-final}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${tag}_full.json 2> gpurun_out/bench_${tag}_full.log; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_${tag}_full.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches_${tag}.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scan -c 1 -f -o gpurun_out/prof_${tag}_scan python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_full_${tag}.log 2>&1; echo "full capture rc=$?"
