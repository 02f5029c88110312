#!/bin/bash
set -u
tag=${1:-r2zc}; out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_render.py tests/test_gpu_pretrain.py tests/test_gpu_ops.py -q -m gpu -x -k "not rulebook" > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 $out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log; echo "bench exit $?"; grep loop $out/${tag}_bench_c2.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2_b.json 2> $out/${tag}_bench_c2_b.log; grep "device-resident" $out/${tag}_bench_c2_b.log
