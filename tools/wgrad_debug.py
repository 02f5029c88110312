import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ponderv2_b200 import _lib
lib = _lib.load()
lib.pv2_wgrad_umma.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
dev = torch.device("cuda:0")
WS = torch.empty(1 << 26, dtype=torch.uint8, device=dev)
torch.set_printoptions(linewidth=220, precision=1, sci_mode=False)
def run(n, cin, cout, x, dy):
    dw = torch.zeros(cout, 1, cin, device=dev)
    rc = lib.pv2_wgrad_umma(_lib.ptr(x), cin, 0, _lib.ptr(dy), cout, 0, None, _lib.ptr(dw), n, n, cin, cout, 1, _lib.ptr(WS), WS.numel(), _lib.stream_ptr())
    torch.cuda.synchronize()
    return rc, dw[:, 0]
lib.pv2_debug_set_wgrad_dump.argtypes = [C.c_void_p]
DBG = torch.zeros(256, device=dev)
lib.pv2_debug_set_wgrad_dump(C.c_void_p(DBG.data_ptr()))
n, cin, cout = 32, 32, 32
# test 1: single row j=0: dy[0, a] = a+1, x[0, b] = 1 -> dw[a,b] = a+1
for jrow in (0,):
    x = torch.zeros(n, cin, device=dev); dy = torch.zeros(n, cout, device=dev)
    x[jrow] = torch.arange(1, cin + 1, device=dev).float() * 100
    dy[jrow] = torch.arange(1, cout + 1, device=dev).float()
    rc, dw = run(n, cin, cout, x, dy)
    print("row", jrow, "rc", rc, "nonzeros", int((dw != 0).sum()), "expected", cin * cout)
    print(dw[:6, :10])
    print("tmem lanes0-3", DBG[128:192].view(4, 16).cpu())
