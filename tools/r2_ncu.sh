#!/bin/bash
# ncu --set full of ONE warmed-up launch of a kernel inside the sparse-conv microbenchmark, exported to CSV on the box.
# Usage (under gpurun): bash tools/r2_ncu.sh <tag> <kernel-regex> <chans> [dtype] [skip]
set -u
tag=$1; rx=$2; ch=$3; dt=${4:-f32}; skip=${5:-6}
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $out/${tag} \
    python tools/spconv_microbench.py --sizes 100000 --chans $ch --dtype $dt > $out/${tag}_ncu.log 2>&1
echo "ncu $tag exit $?"; tail -3 $out/${tag}_ncu.log
ncu -i $out/${tag}.ncu-rep --page details --csv > $out/${tag}_details.csv 2>/dev/null
ncu -i $out/${tag}.ncu-rep --page raw --csv > $out/${tag}_raw.csv 2>/dev/null
ncu -i $out/${tag}.ncu-rep --page source --csv --print-source sass > $out/${tag}_source.csv 2>/dev/null
sz=$(stat -c %s $out/${tag}.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f $out/${tag}.ncu-rep; echo "dropped ${tag}.ncu-rep ($sz bytes)"; fi
