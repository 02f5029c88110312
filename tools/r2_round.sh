#!/bin/bash
# One GPU-box visit of round 2: conv parity subset, A/B microbench of the new kernels, host-overhead split, bench line.
# Usage (under gpurun, from the repo root): bash tools/r2_round.sh <tag>
set -u
tag=${1:-r2}; out=gpurun_out; mkdir -p $out
rm -f $out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "sparse_conv or full_size or linear or spunet_backbone" > $out/${tag}_pytest_conv.log 2>&1; echo "pytest conv exit $?"
tail -12 $out/${tag}_pytest_conv.log
cat $out/parity_report.jsonl 2>/dev/null
timeout 600 python tools/spconv_microbench.py --sizes 100000 --chans 32,96,256 2>&1 | tee $out/${tag}_micro_new.txt
PV2_GG_BX3=0 PV2_WGRAD_MN=0 timeout 600 python tools/spconv_microbench.py --sizes 100000 --chans 32,96,256 2>&1 | tee $out/${tag}_micro_old.txt
timeout 600 python tools/spconv_microbench.py --levels 2>&1 | tee $out/${tag}_micro_levels.txt
timeout 600 python tools/host_overhead.py --steps 10 > $out/${tag}_host_overhead.json 2> $out/${tag}_host_overhead.log; cat $out/${tag}_host_overhead.json; tail -3 $out/${tag}_host_overhead.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.log; echo "bench exit $?"; cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.log
