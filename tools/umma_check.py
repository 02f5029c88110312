"""Development check: tensor-core gather-GEMM vs the SIMT kernel on the same inputs (run on the GPU box)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ponderv2_b200 import _lib, synth
from ponderv2_b200.spconv.pytorch import build_subm_rulebook

lib = _lib.load()
ARGS = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
        C.c_int, C.c_int, C.c_int, C.c_int]
lib.pv2_spconv_gather_gemm_simt.argtypes = ARGS + [C.c_void_p]
lib.pv2_spconv_gather_gemm_umma.argtypes = ARGS + [C.c_void_p, C.c_size_t, C.c_void_p]
WS = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")


def _simt(*a):
    return lib.pv2_spconv_gather_gemm_simt(*a)


def _umma(*a):
    return lib.pv2_spconv_gather_gemm_umma(*a[:-1], _lib.ptr(WS), WS.numel(), a[-1])
dev = torch.device("cuda:0")


def run(fn, x, w3, bias, nbr, n_out):
    cout, kvol, cin = w3.shape
    y = torch.full((n_out, cout), float("nan"), dtype=x.dtype, device=dev)
    rc = fn(_lib.ptr(x), _lib.ptr(w3), w3.stride(0), w3.stride(1), _lib.ptr(bias), _lib.ptr(nbr), _lib.ptr(y),
            x.shape[0], n_out, cin, cout, kvol, _lib.dtype_code(x.dtype), _lib.stream_ptr())
    torch.cuda.synchronize()
    return rc, y


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    c = synth.indoor_cloud(n, 5)
    ind = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int64), c["grid_coord"]], 1).astype(np.int32)).to(dev)
    rb = build_subm_rulebook(ind, (c["grid_coord"].max(0) + 96).tolist(), 3)
    ident = torch.arange(n, dtype=torch.int32, device=dev).view(1, n)
    print(f"n={n} pairs/N={rb.num_pairs / n:.2f}")
    for dtype in (torch.bfloat16, torch.float32):
        for (cin, cout, nbr) in [(32, 32, rb.nbr), (64, 64, rb.nbr), (96, 96, rb.nbr), (128, 96, rb.nbr),
                                 (256, 256, rb.nbr), (192, 128, ident), (16, 48, rb.nbr)]:
            torch.manual_seed(cin + cout)
            kvol = nbr.shape[0]
            x = torch.randn(n, cin, device=dev).to(dtype)
            w3 = (torch.randn(cout, kvol, cin, device=dev) * 0.05).to(dtype)
            bias = torch.randn(cout, device=dev)
            rc0, y0 = run(_simt, x, w3, bias, nbr, n)
            try:
                rc1, y1 = run(_umma, x, w3, bias, nbr, n)
            except Exception as e:  # noqa: BLE001
                print(f"{dtype} {cin}->{cout} K={kvol}: umma raised {e}")
                return
            ref = torch.zeros(n, cout, dtype=torch.float64, device=dev)
            for k in range(kvol):
                m = nbr[k] >= 0
                ref[m] += x[nbr[k][m].long()].double() @ w3[:, k, :].double().t()
            ref += bias.double()
            sc = ref.abs().max().item()
            e0 = (y0.double() - ref).abs().max().item() / sc
            e1 = (y1.double() - ref).abs().max().item() / sc
            nan1 = int(torch.isnan(y1.float()).sum().item())
            # timing
            for fn, tag in ((_simt, "simt"), (_umma, "umma")):
                run(fn, x, w3, bias, nbr, n)
                t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(5):
                    y = torch.empty((n, cout), dtype=x.dtype, device=dev)
                    fn(_lib.ptr(x), _lib.ptr(w3), w3.stride(0), w3.stride(1), _lib.ptr(bias), _lib.ptr(nbr), _lib.ptr(y),
                       n, n, cin, cout, kvol, _lib.dtype_code(x.dtype), _lib.stream_ptr())
                t1.record(); torch.cuda.synchronize()
                if tag == "simt":
                    ms0 = t0.elapsed_time(t1) / 5
                else:
                    ms1 = t0.elapsed_time(t1) / 5
            flops = 2.0 * (nbr >= 0).sum().item() * cin * cout
            print(f"{str(dtype):15s} {cin:3d}->{cout:3d} K={kvol:2d} rc=({rc0},{rc1}) err simt {e0:.2e} umma {e1:.2e} "
                  f"nan {nan1} | simt {ms0:.3f} ms umma {ms1:.3f} ms ({flops / ms1 * 1e-9:.1f} TFLOP/s useful)")


if __name__ == "__main__":
    main()
