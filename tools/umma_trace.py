"""Development: per-chunk timeline of CTA 0 of the tensor-core gather-GEMM (clock64 stamps)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ponderv2_b200 import _lib, synth
from ponderv2_b200.spconv.pytorch import build_subm_rulebook
lib = _lib.load()
ARGS = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
lib.pv2_spconv_gather_gemm_umma.argtypes = ARGS
lib.pv2_debug_set_trace.argtypes = [C.c_void_p]
dev = torch.device("cuda:0")
n = 100000
c = synth.indoor_cloud(n, 5)
ind = torch.from_numpy(np.concatenate([np.zeros((n, 1), np.int64), c["grid_coord"]], 1).astype(np.int32)).to(dev)
rb = build_subm_rulebook(ind, (c["grid_coord"].max(0) + 96).tolist(), 3)
WS = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for dtype, cin, cout in [(torch.bfloat16, 32, 32), (torch.bfloat16, 96, 96), (torch.float32, 96, 96)]:
    x = torch.randn(n, cin, device=dev).to(dtype); w3 = (torch.randn(cout, 27, cin, device=dev) * 0.05).to(dtype)
    y = torch.empty(n, cout, dtype=dtype, device=dev)
    tr = torch.zeros(4096, dtype=torch.int64, device=dev)
    for rep in range(2):
        lib.pv2_debug_set_trace(C.c_void_p(tr.data_ptr()) if rep == 1 else None)
        lib.pv2_spconv_gather_gemm_umma(_lib.ptr(x), _lib.ptr(w3), w3.stride(0), w3.stride(1), None, _lib.ptr(rb.nbr), _lib.ptr(y),
                                        n, n, cin, cout, 27, _lib.dtype_code(dtype), _lib.ptr(WS), WS.numel(), _lib.stream_ptr())
        torch.cuda.synchronize()
    lib.pv2_debug_set_trace(None)
    t = tr.cpu().numpy().reshape(-1, 8)
    t = t[t[:, 5] > 0]
    t0 = t[:, [0, 5]].min()
    print(f"--- {dtype} {cin}->{cout}: {len(t)} chunks; columns: p.wait_empty_begin p.wait_empty_end p.issued p.landed p.arrived | m.wait_begin m.woke m.issued (cycles from start)")
    for i, row in enumerate(t[:12]):
        print(i, " ".join(f"{int(v - t0):7d}" if v > 0 else "      -" for v in row))
    print("last", " ".join(f"{int(v - t0):7d}" if v > 0 else "      -" for v in t[-1]))
