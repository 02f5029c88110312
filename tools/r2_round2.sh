#!/bin/bash
# Full GPU suite + the three bench workloads (+ 2-GPU c2 when two devices are visible).
set -u
tag=${1:-r2f}; out=gpurun_out; mkdir -p $out
rm -f $out/parity_report.jsonl
timeout 1500 python -m pytest tests -q -m gpu -x > $out/${tag}_pytest_all.log 2>&1; echo "pytest all exit $?"; tail -12 $out/${tag}_pytest_all.log
cp $out/parity_report.jsonl $out/${tag}_parity_report.jsonl 2>/dev/null
for w in c2 c3 c4; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $w $( [ $w = c2 ] || echo --no-cpu-baseline ) > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.log; echo "bench $w exit $?"; cat $out/${tag}_bench_$w.json; tail -4 $out/${tag}_bench_$w.log
done
n=$(nvidia-smi -L | wc -l)
if [ "$n" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/${tag}_bench_c2_2gpu.json 2> $out/${tag}_bench_c2_2gpu.log; echo "bench 2gpu exit $?"; cat $out/${tag}_bench_c2_2gpu.json; tail -4 $out/${tag}_bench_c2_2gpu.log
fi
