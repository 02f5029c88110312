#!/bin/bash
# Full GPU suite (log to file), C2 bench, host enqueue time + cProfile of the step.
set -u
tag=${1:-r2l}; out=gpurun_out; mkdir -p $out
rm -f $out/parity_report.jsonl
timeout 1800 python -m pytest tests -q -m gpu > $out/${tag}_pytest_all.log 2>&1; echo "pytest all exit $?"; tail -15 $out/${tag}_pytest_all.log
cp $out/parity_report.jsonl $out/${tag}_parity_report.jsonl 2>/dev/null
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log; echo "bench exit $?"; cut -c1-300 $out/${tag}_bench_c2.json; tail -4 $out/${tag}_bench_c2.log
timeout 600 python tools/host_overhead.py --steps 10 > $out/${tag}_host_overhead.json 2> $out/${tag}_host_overhead.log; cat $out/${tag}_host_overhead.json
timeout 600 python tools/host_profile.py > $out/${tag}_host_profile.txt 2>&1; head -60 $out/${tag}_host_profile.txt
