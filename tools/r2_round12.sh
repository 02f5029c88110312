#!/bin/bash
# 1-GPU: full suite, smoke(), bench c2/c3/c4, host overhead.
set -u
tag=${1:-r2t}; out=gpurun_out; mkdir -p $out
rm -f $out/parity_report.jsonl
timeout 1800 python -m pytest tests -q -m gpu > $out/${tag}_pytest_all.log 2>&1; echo "pytest all exit $?"; tail -6 $out/${tag}_pytest_all.log
cp $out/parity_report.jsonl $out/${tag}_parity_report.jsonl 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $out/${tag}_smoke.log
for w in c2 c3 c4; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $w > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.log; echo "bench $w exit $?"; grep loop $out/${tag}_bench_$w.log
done
timeout 600 python tools/host_overhead.py --steps 10 > $out/${tag}_host_overhead.json 2> $out/${tag}_host_overhead.log; grep -A12 host_enqueue_ms $out/${tag}_host_overhead.json | tr -d '\n '; echo
