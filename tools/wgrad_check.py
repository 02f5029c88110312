"""Development check: tensor-core wgrad vs fp64 reference."""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ponderv2_b200 import _lib, synth
from ponderv2_b200.spconv.pytorch import build_subm_rulebook
lib = _lib.load()
lib.pv2_wgrad_umma.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
lib.pv2_spconv_wgrad_simt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p]
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
c = synth.indoor_cloud(n, 5)
ind = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int64), c["grid_coord"]], 1).astype(np.int32)).to(dev)
rb = build_subm_rulebook(ind, (c["grid_coord"].max(0) + 96).tolist(), 3)
WS = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for cin, cout, nbr in [(32, 32, rb.nbr), (96, 96, rb.nbr), (64, 128, rb.nbr), (256, 256, rb.nbr), (128, 64, None), (192, 68, None)]:
    kvol = 27 if nbr is not None else 1
    torch.manual_seed(cin)
    x = torch.randn(n, cin, device=dev); dy = torch.randn(n, cout, device=dev)
    ref = torch.zeros(cout, kvol, cin, dtype=torch.float64, device=dev)
    for k in range(kvol):
        if nbr is None:
            ref[:, k] = dy.double().t() @ x.double()
        else:
            m = nbr[k] >= 0
            ref[:, k] = dy[m].double().t() @ x[nbr[k][m].long()].double()
    res = {}
    for tag in ("simt", "umma"):
        nb = nbr if nbr is not None else torch.arange(n, dtype=torch.int32, device=dev).view(1, n)
        def call():
            dw = torch.zeros(cout, kvol, cin, device=dev)
            if tag == "simt":
                rc = lib.pv2_spconv_wgrad_simt(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(nb), _lib.ptr(dw), n, n, cin, cout, kvol, 0, _lib.stream_ptr())
            else:
                rc = lib.pv2_wgrad_umma(_lib.ptr(x), cin, 0, _lib.ptr(dy), cout, 0, _lib.ptr(nbr) if nbr is not None else None, _lib.ptr(dw),
                                        n, n, cin, cout, kvol, _lib.ptr(WS), WS.numel(), _lib.stream_ptr())
            return rc, dw
        rc, dw = call(); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3): call()
        t1.record(); torch.cuda.synchronize()
        res[tag] = (rc, (dw.double() - ref).abs().max().item() / ref.abs().max().item(), t0.elapsed_time(t1) / 3)
    print(f"{cin:3d}->{cout:3d} K={kvol:2d}: simt rc={res['simt'][0]} err {res['simt'][1]:.2e} {res['simt'][2]:.3f} ms | "
          f"umma rc={res['umma'][0]} err {res['umma'][1]:.2e} {res['umma'][2]:.3f} ms")
