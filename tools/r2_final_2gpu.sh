#!/bin/bash
# 2-GPU: c2 / c3 / c4 under torchrun (NCCL, overlapped all-reduce).
set -u
tag=${1:-r2zb}; out=gpurun_out; mkdir -p $out
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 "$@" > $out/${tag}_bench_${name}_2gpu.json 2> $out/${tag}_bench_${name}_2gpu.log; echo "$name exit $?"; grep -h "loop" $out/${tag}_bench_${name}_2gpu.log | sort -u | cut -c1-200; }
run c2 --workload c2
run c3 --workload c3
run c4 --workload c4
