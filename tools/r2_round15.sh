#!/bin/bash
# in-run A/B of the metadata-buffer rule (same box): bench c2 + micro levels, default vs PV2_GG_META_BUFS=3 vs 2
set -u
tag=${1:-r2y}; out=gpurun_out; mkdir -p $out
for mb in 0 3 2 0 3; do
  PV2_GG_META_BUFS=$mb timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2_mb$mb.json 2> $out/${tag}_bench_c2_mb$mb.log; echo "mb=$mb: $(grep 'device-resident' $out/${tag}_bench_c2_mb$mb.log)"
done
for mb in 0 3; do
  PV2_GG_META_BUFS=$mb timeout 600 python tools/spconv_microbench.py --levels > $out/${tag}_micro_levels_mb$mb.txt 2>&1
  echo "mb=$mb"; grep -A3 "^L0\|^L1" $out/${tag}_micro_levels_mb$mb.txt | cut -c1-110
done
