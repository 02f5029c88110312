"""Sparse-conv microbenchmark (BASELINE.json configs[4]): rulebook build + SubMConv3d
gather-GEMM fwd / dgrad / wgrad per (N, Cin, Cout), CUDA-event timed, reported as algorithmic GB/s against the measured
HBM peak and as issued TFLOP/s.  Run on the GPU box:

    python tools/spconv_microbench.py [--sizes 10000,100000,1000000] [--chans 32,64,128,256] [--dtype f32|bf16]
                                      [--levels]   # instead: the five U-Net levels of the C2 scene with App. C's shapes
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ponderv2_b200 import _lib, synth  # noqa: E402
from ponderv2_b200.spconv import pytorch as sp  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


def indices_for(n, seed):
    c = synth.indoor_cloud(n, seed)
    ind = np.concatenate([np.zeros((n, 1), np.int64), c["grid_coord"]], 1).astype(np.int32)
    return torch.from_numpy(ind).to(DEV), (c["grid_coord"].max(0) + 96).tolist()


def bench_layer(rb, n, cin, cout, dtype, hbm):
    nbr = rb.nbr
    K = nbr.shape[0]
    b = 4 if dtype == torch.float32 else 2
    x = torch.randn(n, cin, device=DEV).to(dtype)
    w3 = (torch.randn(cout, K, cin, device=DEV) * 0.02).to(dtype)
    dy = torch.randn(n, cout, device=DEV).to(dtype)
    wt = w3.flip(1).permute(2, 1, 0).contiguous()
    P = rb.num_pairs
    t_f = timed(lambda: sp._gather_gemm(x, w3, None, rb.tmap, n))
    t_d = timed(lambda: sp._gather_gemm(dy, wt, None, rb.tmap, n))
    t_w = timed(lambda: sp._wgrad(x, dy, rb.tmap, K))
    by_f = n * cin * b + n * cout * b + K * cin * cout * b + 4 * K * n
    by_w = n * (cin + cout) * b + K * cin * cout * 4 + 4 * K * n
    fl = 2.0 * P * cin * cout
    return dict(n=n, cin=cin, cout=cout, K=K, pairs=P, fwd_us=t_f, dgrad_us=t_d, wgrad_us=t_w,
                fwd_gbs=by_f / t_f * 1e-3, fwd_frac=by_f / t_f * 1e-3 / hbm, wgrad_gbs=by_w / t_w * 1e-3,
                fwd_tflops=fl / t_f * 1e-6, wgrad_tflops=fl / t_w * 1e-6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="10000,100000,1000000")
    ap.add_argument("--chans", default="32,64,128,256")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--levels", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    hbm = peaks()
    _lib.load()
    rows = []
    if args.levels:
        ind, shape = indices_for(100_000, 2000)
        shapes = [[(32, 32), (128, 96), (96, 96)], [(32, 32), (128, 96), (96, 96)], [(64, 64), (192, 128), (128, 128)],
                  [(128, 128), (384, 256), (256, 256)], [(256, 256)]]
        for lvl in range(5):
            n = ind.shape[0]
            t_rb = timed(lambda: sp.build_subm_rulebook(ind, shape, 3), iters=10)
            rb = sp.build_subm_rulebook(ind, shape, 3)
            print(f"L{lvl}: N={n} pairs={rb.num_pairs} rulebook {t_rb:.1f} us", flush=True)
            for cin, cout in shapes[lvl]:
                r = bench_layer(rb, n, cin, cout, dtype, hbm)
                r["level"] = lvl
                rows.append(r)
                print(f"   {cin:3d}->{cout:3d}  fwd {r['fwd_us']:8.1f} us ({r['fwd_gbs']:7.1f} GB/s alg, "
                      f"{r['fwd_tflops']:6.1f} TF/s)  dgrad {r['dgrad_us']:8.1f}  wgrad {r['wgrad_us']:8.1f} us "
                      f"({r['wgrad_tflops']:6.1f} TF/s)", flush=True)
            if lvl < 4:
                t_dn = timed(lambda: sp.build_down_rulebook(ind, shape), iters=10)
                d = sp.build_down_rulebook(ind, shape)
                print(f"   down rulebook {t_dn:.1f} us -> {d.out_indices.shape[0]}")
                ind, shape = d.out_indices.contiguous(), d.out_shape
    else:
        for n in [int(s) for s in args.sizes.split(",")]:
            ind, shape = indices_for(n, 5000 + n % 997)
            t_rb = timed(lambda: sp.build_subm_rulebook(ind, shape, 3), iters=10)
            rb = sp.build_subm_rulebook(ind, shape, 3)
            rb_bytes = 16 * n + 4 * 27 * n
            print(f"N={n} pairs={rb.num_pairs} ({rb.num_pairs / n:.2f}/voxel) rulebook {t_rb:.1f} us "
                  f"({rb_bytes / t_rb * 1e-3:.1f} GB/s alg = {rb_bytes / t_rb * 1e-3 / hbm:.3f} of {hbm:.0f})", flush=True)
            rows.append(dict(n=n, rulebook_us=t_rb, rulebook_gbs=rb_bytes / t_rb * 1e-3))
            for c in [int(s) for s in args.chans.split(",")]:
                r = bench_layer(rb, n, c, c, dtype, hbm)
                rows.append(r)
                print(f"   C={c:3d}  fwd {r['fwd_us']:8.1f} us  {r['fwd_gbs']:7.1f} GB/s alg = {r['fwd_frac']:.3f} of HBM "
                      f"peak, {r['fwd_tflops']:6.1f} TF/s | dgrad {r['dgrad_us']:8.1f} us | wgrad {r['wgrad_us']:8.1f} us "
                      f"{r['wgrad_gbs']:7.1f} GB/s {r['wgrad_tflops']:6.1f} TF/s", flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps(dict(dtype=args.dtype, hbm_peak_gbs=hbm, rows=rows), indent=1))


if __name__ == "__main__":
    main()
