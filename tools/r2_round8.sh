#!/bin/bash
# 1-GPU: full suite + c3/c4 bench on the current tree.
set -u
tag=${1:-r2n}; out=gpurun_out; mkdir -p $out
rm -f $out/parity_report.jsonl
timeout 1800 python -m pytest tests -q -m gpu > $out/${tag}_pytest_all.log 2>&1; echo "pytest all exit $?"; tail -8 $out/${tag}_pytest_all.log
cp $out/parity_report.jsonl $out/${tag}_parity_report.jsonl 2>/dev/null
for w in c3 c4; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $w > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.log; echo "bench $w exit $?"; tail -3 $out/${tag}_bench_$w.log
done
