#!/bin/bash
# 1-GPU: split-K bf16x3 on the deep levels: conv parity tests, micro levels at several caps, bench c2.
set -u
tag=${1:-r2v}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "full_size or sparse_conv or backbone or linear or v1m3" > $out/${tag}_pytest_conv.log 2>&1; echo "pytest exit $?"; tail -5 $out/${tag}_pytest_conv.log
for cap in 0 16; do
  PV2_GG_KSPLIT_MAX=$cap timeout 600 python tools/spconv_microbench.py --levels > $out/${tag}_micro_levels_cap$cap.txt 2>&1; echo "micro cap $cap exit $?"
  grep -A3 "^L2\|^L3\|^L4" $out/${tag}_micro_levels_cap$cap.txt | cut -c1-120
done
PV2_GG_BX3_SPLIT=0 timeout 600 python tools/spconv_microbench.py --levels > $out/${tag}_micro_levels_old.txt 2>&1
grep -A3 "^L2\|^L3\|^L4" $out/${tag}_micro_levels_old.txt | cut -c1-120
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log; echo "bench exit $?"; grep loop $out/${tag}_bench_c2.log
