#!/bin/bash
set -u
tag=${1:-r2g}; out=gpurun_out; mkdir -p $out
python tools/bf16_grad_diag.py 2>&1 | tail -70 | tee $out/${tag}_bf16_diag.txt
timeout 900 python -m pytest tests/test_gpu_pretrain.py -q -x 2>&1 | tail -8
for cap in 0 8 16; do echo "== ksplit cap $cap"; PV2_GG_KSPLIT_MAX=$cap timeout 600 python tools/spconv_microbench.py --levels 2>&1 | grep -E "^L[234]|->" | tee -a $out/${tag}_ksplit.txt; done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log; echo "bench exit $?"; cut -c1-400 $out/${tag}_bench_c2.json; tail -4 $out/${tag}_bench_c2.log
bash tools/r2_ncu_step.sh $tag "umma_gather_gemm_persistent_kernel<.bool.1" umma_wgrad_mn umma_gather_gemm_kernel field_post_fwd field_post_bwd field_sample_fwd field_sample_bwd ray_composite ray_resample probe_subm scatter_mean bn_apply_fwd
