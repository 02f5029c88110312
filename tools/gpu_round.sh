#!/bin/bash
# One GPU-box visit: parity tests, microbench, the bench line, the ncu launch list of one warmed-up step and full captures
# of the top kernels (exported to CSV on the box; the .ncu-rep files are dropped when large so gpurun_out stays < 64 MiB).
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [kernel-regex ...]
set -u
tag=${1:-r1}; shift || true
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?"
tail -15 $out/${tag}_pytest.log
if [ "${SKIP_MICRO:-0}" != "1" ]; then
  timeout 600 python tools/spconv_microbench.py --levels --out $out/${tag}_micro_levels.json 2>&1 | tee $out/${tag}_micro_levels.txt
  if [ "${MICRO_AB:-0}" = "1" ]; then
    env ${MICRO_AB_ENV:-PV2_GG_GROUPS=2} timeout 600 python tools/spconv_microbench.py --levels 2>&1 | tee $out/${tag}_micro_levels_ab.txt
  fi
  if [ "${MICRO_NOORDER:-0}" = "1" ]; then
    PV2_ROW_ORDER=0 timeout 600 python tools/spconv_microbench.py --levels 2>&1 | tee $out/${tag}_micro_levels_noorder.txt
  fi
  timeout 600 python tools/spconv_microbench.py --sizes 100000,1000000 --out $out/${tag}_micro_c5.json 2>&1 | tee $out/${tag}_micro_c5.txt
fi
if [ "${NCU_MICRO:-0}" = "1" ]; then
  # one warmed-up launch of each conv kernel at the C2 level-0 decoder shape (100 k voxels, 96 -> 96), full metric set +
  # per-instruction stall samples (SASS)
  for kn in umma_gather_gemm umma_wgrad; do
    timeout 600 ncu --set full --clock-control none -k regex:$kn -s 6 -c 1 -f -o $out/${tag}_micro_$kn \
        python tools/spconv_microbench.py --sizes 100000 --chans 96 > $out/${tag}_ncu_micro_$kn.log 2>&1
    echo "ncu micro $kn exit $?"
    ncu -i $out/${tag}_micro_$kn.ncu-rep --page details --csv > $out/${tag}_micro_${kn}_details.csv 2>/dev/null
    ncu -i $out/${tag}_micro_$kn.ncu-rep --page raw --csv > $out/${tag}_micro_${kn}_raw.csv 2>/dev/null
    ncu -i $out/${tag}_micro_$kn.ncu-rep --page source --csv --print-source sass > $out/${tag}_micro_${kn}_source.csv 2>/dev/null
  done
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.log; echo "bench exit $?"
cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file $out/${tag}_launches.csv python bench.py --profile-step > $out/${tag}_ncu_list.log 2>&1
echo "ncu list exit $?"
python tools/launch_summary.py $out/${tag}_launches.csv 45 | tee $out/${tag}_launch_summary.txt
i=0
for rx in "$@"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:$rx -c ${NCU_COUNT:-6} \
      -f -o $out/${tag}_full_$i python bench.py --profile-step > $out/${tag}_ncu_full_$i.log 2>&1
  echo "ncu full $rx exit $?"
  ncu -i $out/${tag}_full_$i.ncu-rep --page raw --csv > $out/${tag}_full_${i}_raw.csv 2>/dev/null
  ncu -i $out/${tag}_full_$i.ncu-rep --page details --csv > $out/${tag}_full_${i}_details.csv 2>/dev/null
  sz=$(stat -c %s $out/${tag}_full_$i.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 12000000 ]; then rm -f $out/${tag}_full_$i.ncu-rep; echo "dropped ${tag}_full_$i.ncu-rep ($sz bytes)"; fi
done
du -sh $out
