#!/bin/bash
# 2-GPU round: new PDNorm tests, 1-GPU bench (clock sampler moved), 2-GPU bench over NCCL (overlapped all-reduce).
set -u
tag=${1:-r2m}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "pdnorm or v1m3" > $out/${tag}_pytest_pdnorm.log 2>&1; echo "pytest exit $?"; tail -15 $out/${tag}_pytest_pdnorm.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log; echo "bench1 exit $?"; tail -4 $out/${tag}_bench_c2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $out/${tag}_bench_c2_2gpu.json 2> $out/${tag}_bench_c2_2gpu.log; echo "bench2 exit $?"; cut -c1-400 $out/${tag}_bench_c2_2gpu.json; tail -12 $out/${tag}_bench_c2_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload c3 > $out/${tag}_bench_c3_2gpu.json 2> $out/${tag}_bench_c3_2gpu.log; echo "bench3 exit $?"; cut -c1-300 $out/${tag}_bench_c3_2gpu.json; tail -5 $out/${tag}_bench_c3_2gpu.log
