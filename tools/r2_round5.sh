#!/bin/bash
set -u
tag=${1:-r2j}; out=gpurun_out; mkdir -p $out
rm -f $out/parity_report.jsonl
timeout 1200 python -m pytest tests/test_gpu_render.py tests/test_gpu_pretrain.py -q > $out/${tag}_pytest_render.log 2>&1; echo "pytest exit $?"; tail -15 $out/${tag}_pytest_render.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "bn_act" > $out/${tag}_pytest_bn.log 2>&1; echo "pytest bn exit $?"; tail -5 $out/${tag}_pytest_bn.log
cp $out/parity_report.jsonl $out/${tag}_parity_report.jsonl 2>/dev/null
for w in c2 c4; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.log; echo "bench $w exit $?"; cut -c1-260 $out/${tag}_bench_$w.json; tail -4 $out/${tag}_bench_$w.log
done
