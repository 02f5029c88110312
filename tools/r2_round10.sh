#!/bin/bash
# 1-GPU: producer-group A/B of the persistent fp32 gather-GEMM (micro levels), then bench c2 per setting.
set -u
tag=${1:-r2p}; out=gpurun_out; mkdir -p $out
for g in 2 3 4; do
  PV2_GG_GROUPS=$g timeout 600 python tools/spconv_microbench.py --levels > $out/${tag}_micro_levels_g$g.txt 2>&1; echo "micro g$g exit $?"
  grep -A3 "^L0\|^L1" $out/${tag}_micro_levels_g$g.txt | cut -c1-120
done
for g in 2 3 4; do
  PV2_GG_GROUPS=$g timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2_g$g.json 2> $out/${tag}_bench_c2_g$g.log; echo "bench g$g exit $?"; grep "loop" $out/${tag}_bench_c2_g$g.log
done
PV2_GG_GROUPS=3 timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "full_size or sparse_conv or backbone or linear" > $out/${tag}_pytest_g3.log 2>&1; echo "pytest g3 exit $?"; tail -3 $out/${tag}_pytest_g3.log
