#!/bin/bash
# 1-GPU: render + pretrain tests, bench c2, launch list of one step.
set -u
tag=${1:-r2u}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_pretrain.py -q -m gpu > $out/${tag}_pytest_render.log 2>&1; echo "pytest exit $?"; tail -3 $out/${tag}_pytest_render.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log; echo "bench exit $?"; grep loop $out/${tag}_bench_c2.log
bash tools/r2_ncu_step.sh $tag
