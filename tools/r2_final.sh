#!/bin/bash
# Final 1-GPU evidence round: default bench line (with CPU baseline), UNet3D-projection variant, c3, c4, host overhead,
# ncu launch list + traffic + a few --set full captures.
set -u
tag=${1:-r2z}; out=gpurun_out; mkdir -p $out
timeout 900 python bench.py --steps 20 --warmup 3 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log; echo "bench c2 exit $?"; grep "loop\|cpu baseline done" $out/${tag}_bench_c2.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 3 --projection unet3d > $out/${tag}_bench_c2_unet3d.json 2> $out/${tag}_bench_c2_unet3d.log; echo "bench c2 unet3d exit $?"; grep loop $out/${tag}_bench_c2_unet3d.log
for w in c3 c4; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload $w > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.log; echo "bench $w exit $?"; grep loop $out/${tag}_bench_$w.log
done
timeout 600 python tools/host_overhead.py --steps 10 > $out/${tag}_host_overhead.json 2> $out/${tag}_host_overhead.log
bash tools/r2_ncu_step.sh $tag "umma_gather_gemm_persistent_kernel<1, 2, 3, 1>"
