#!/bin/bash
# 2-GPU: c3 with / without all-reduce overlap, NCCL logs kept.
set -u
tag=${1:-r2o}; out=gpurun_out; mkdir -p $out
env | grep -i nccl
run() { name=$1; shift; PV2_NCCL_LOG_COPY=$out/${tag}_nccl_$name timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.log; echo "$name exit $?"; grep -h "loop\|nccl:" $out/${tag}_bench_$name.log | sort -u | cut -c1-900; }
run c3_overlap --workload c3
run c3_nooverlap --workload c3 --no-overlap
run c2_nooverlap --workload c2 --no-overlap
run c4_overlap --workload c4
ls $out/${tag}_nccl_c3_overlap | head; for f in $out/${tag}_nccl_*/*.log; do head -c 20000 $f > $f.head; rm $f; done
