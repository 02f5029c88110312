#!/bin/bash
# bf16 path: parity tests, then the C5 microbenchmark in bf16 with the TMA gather4 kernel on and off.
set -u
tag=${1:-r2e}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "bf16" > $out/${tag}_pytest_bf16.log 2>&1; echo "pytest bf16 exit $?"; tail -8 $out/${tag}_pytest_bf16.log
timeout 600 python tools/spconv_microbench.py --dtype bf16 --sizes 100000,1000000 --chans 64,128 2>&1 | tee $out/${tag}_micro_bf16_tma.txt
PV2_GG_TMA=0 timeout 600 python tools/spconv_microbench.py --dtype bf16 --sizes 100000,1000000 --chans 64,128 2>&1 | tee $out/${tag}_micro_bf16_cpasync.txt
timeout 1200 python -m pytest tests -q -m gpu > $out/${tag}_pytest_all.log 2>&1; echo "pytest all exit $?"; tail -15 $out/${tag}_pytest_all.log
