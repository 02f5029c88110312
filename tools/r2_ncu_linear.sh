#!/bin/bash
# ncu --set full of the 7th render-linear launch of a warmed-up C2 step (72 -> 128 backward layer, read-modify-write epilogue)
out=gpurun_out; mkdir -p $out; tag=r2zf_linear_act3
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:umma_gather_gemm_persistent -s 24 -c 1 -f -o $out/$tag python bench.py --profile-step > $out/${tag}_ncu.log 2>&1
echo "ncu exit $?"; tail -2 $out/${tag}_ncu.log
ncu -i $out/$tag.ncu-rep --page details --csv > $out/${tag}_details.csv 2>/dev/null
ncu -i $out/$tag.ncu-rep --page source --csv --print-source sass > $out/${tag}_source.csv 2>/dev/null
rm -f $out/$tag.ncu-rep
python tools/ncu_brief.py $out/${tag}_details.csv | head -24
