#!/bin/bash
# ncu evidence of one warmed-up C2 step: the launch list (time + DRAM bytes per launch) and --set full captures of the
# kernels SURVEY 8(d) names, each exported to CSV on the box (reports dropped when large).
# Usage (under gpurun): bash tools/r2_ncu_step.sh <tag> [regex ...]
set -u
tag=$1; shift
out=gpurun_out; mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file $out/${tag}_launches.csv python bench.py --profile-step > $out/${tag}_ncu_list.log 2>&1
echo "ncu list exit $?"
python tools/launch_summary.py $out/${tag}_launches.csv 50 | tee $out/${tag}_launch_summary.txt
python tools/traffic_from_launches.py $out/${tag}_launches.csv $out/${tag}_traffic.json | tail -8
i=0
for rx in "$@"; do
  i=$((i+1))
  name=$(echo $rx | tr -c 'a-zA-Z0-9_' '_' | cut -c1-40)
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx -c 1 \
      -f -o $out/${tag}_full_$name python bench.py --profile-step > $out/${tag}_ncu_full_$name.log 2>&1
  echo "ncu full $rx exit $?"
  ncu -i $out/${tag}_full_$name.ncu-rep --page details --csv > $out/${tag}_full_${name}_details.csv 2>/dev/null
  python tools/ncu_brief.py $out/${tag}_full_${name}_details.csv
  rm -f $out/${tag}_full_$name.ncu-rep
done
du -sh $out
