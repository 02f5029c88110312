#!/bin/bash
# 1-GPU: metadata-ahead persistent kernel: micro levels, conv parity tests, bench c2.
set -u
tag=${1:-r2r}; out=gpurun_out; mkdir -p $out
timeout 600 python tools/spconv_microbench.py --levels > $out/${tag}_micro_levels.txt 2>&1; echo "micro exit $?"
grep -A3 "^L0\|^L1" $out/${tag}_micro_levels.txt | cut -c1-120
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "full_size or sparse_conv or backbone or linear" > $out/${tag}_pytest_conv.log 2>&1; echo "pytest exit $?"; tail -3 $out/${tag}_pytest_conv.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log; echo "bench exit $?"; grep "loop" $out/${tag}_bench_c2.log
