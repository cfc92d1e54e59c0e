#!/bin/bash
# gpurun_out/<tag>_* (tools/final_measure.sh) -> profiles/<tag>_*: the summaries that are committed.  usage: tools/export_profiles.sh <tag>
tag=${1:-r02}
out=profiles
exp() { # report, kernel regex, name
    rep=gpurun_out/${tag}_$1.ncu-rep; [ -f "$rep" ] || return
    ncu -i "$rep" --page raw --csv 2>/dev/null > /tmp/${tag}_$1_raw.csv && python tools/ncu_summary.py /tmp/${tag}_$1_raw.csv > $out/${tag}_ncu_full_$1.txt
}
exp k_scan; exp xa; exp inflate
for k in k_scan k_xa; do
    rep=gpurun_out/${tag}_$([ $k = k_scan ] && echo k_scan || echo xa).ncu-rep; [ -f "$rep" ] || continue
    ncu -i "$rep" --page source --csv --print-source cuda,sass --kernel-name regex:$k --launch-count 1 2>/dev/null > /tmp/${tag}_${k}_cs.csv && python tools/ncu_lines.py /tmp/${tag}_${k}_cs.csv 60 | cut -c1-220 > $out/${tag}_${k}_source_lines.txt
done
[ -f gpurun_out/${tag}_launches.csv ] && cp gpurun_out/${tag}_launches.csv $out/${tag}_launches.csv
for f in bench_full; do [ -f gpurun_out/${tag}_$f.json ] && cp gpurun_out/${tag}_$f.json $out/${tag}_$f.json; done
# what the SASS of the built library shows (Blackwell-native byte-stream kernels: bulk copies, mbarriers, 128-bit CAS; no tensor-core path here)
{
    echo "# cuobjdump -sass iteres_b200/csrc/libiteres_gpu.so: occurrences per kernel"
    cuobjdump -sass iteres_b200/csrc/libiteres_gpu.so | awk '/Function : /{f=$3} /UBLKCP/{a[f" UBLKCP (cp.async.bulk: TMA 1-D bulk copy)"]++} /UBLKPF/{a[f" UBLKPF (cp.async.bulk.prefetch.L2)"]++} /SYNCS/{a[f" SYNCS (mbarrier)"]++} /ATOMG.*128/{a[f" ATOMG.E.CAS.128"]++} /LDGSTS/{a[f" LDGSTS (cp.async)"]++} /REDUX/{a[f" REDUX (redux.sync)"]++} /RED\.E/{a[f" RED (red.global)"]++} END{for(k in a) print a[k], k}' | sort -k2
    echo "# md5 per kernel (tools/sass_md5.sh)"
    tools/sass_md5.sh
} > $out/${tag}_sass.txt
ls -la $out | grep ${tag}_
[ -f $out/${tag}_launches.csv ] && python tools/launch_shares.py $out/${tag}_launches.csv > $out/${tag}_launch_shares.txt
