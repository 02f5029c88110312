#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of one warmed-up step and full captures of the
# top kernels.  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [kernel-regex ...]
set -u
tag=${1:-r1}; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.log; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --profile-step > gpurun_out/${tag}_ncu_list.log 2>&1
echo "ncu list exit $?"
python tools/launch_summary.py gpurun_out/${tag}_launches.csv 40 | tee gpurun_out/${tag}_launch_summary.txt
i=0
for rx in "$@"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx -c ${NCU_COUNT:-12} \
      -f -o gpurun_out/${tag}_full_$i python bench.py --profile-step > gpurun_out/${tag}_ncu_full_$i.log 2>&1
  echo "ncu full $rx exit $?"
done
