"""ctypes binding of libpv2_b200.so (the C ABI declared in include/pv2_b200.h).

There is no CPU fallback: if the library is missing, or a call is made without a CUDA tensor,
the failure is loud (RuntimeError).  Loading the library itself needs no GPU, so the symbol
table can be checked on a CPU-only box.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libpv2_b200.so"
_lib = None

_i32p = C.c_void_p
_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/pv2_b200.h
SIGNATURES = {
    "pv2_version": (_int, []),
    "pv2_error_string": (C.c_char_p, [_int]),
    "pv2_sm_count": (_int, []),
    "pv2_launch_count": (_i64, []),
    "pv2_set_option": (_int, [C.c_char_p, _int]),
    "pv2_get_option": (_int, [C.c_char_p]),
    "pv2_rulebook_workspace_bytes": (_sz, [_i64]),
    "pv2_rulebook_subm": (_int, [_vp, _i64, C.POINTER(C.c_int32), _int, _vp, _vp, _vp, _sz, _vp]),
    "pv2_rulebook_down": (_int, [_vp, _i64, C.POINTER(C.c_int32), _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pv2_rulebook_down_maps": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "pv2_make_indices": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "pv2_rulebook_row_order_workspace_bytes": (_sz, [_i64]),
    "pv2_rulebook_row_order": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pv2_spconv_gather_gemm": (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _int, _int, _vp,
                                       _sz, _vp]),
    "pv2_spconv_dgrad_weights": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _vp]),
    "pv2_spconv_workspace_bytes": (_sz, [_i64, _int, _int, _int, _int]),
    "pv2_spconv_wgrad": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _int, _int, _vp, _sz, _vp]),
    "pv2_wgrad_workspace_bytes": (_sz, [_i64, _i64, _int, _int]),
    "pv2_linear_workspace_bytes": (_sz, [_i64, _int, _int, _int]),
    "pv2_linear": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _i64, _i64, _int, _int, _vp, _i64, _i64, _i64, _int, _int,
                           _vp, _sz, _vp]),
    "pv2_dense_wgrad": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _int, _int, _vp, _vp, _sz, _vp]),
    "pv2_field_sample_fwd": (_int, [_vp, _vp, _i64, _int, _int, _int, _int, _int, _int, _vp, _i64, _i64, _vp, _i64, _vp]),
    "pv2_field_post_fwd": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _int, _int, _int, _int, _vp,
                                   _vp, _vp]),
    "pv2_field_post_bwd": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int,
                                   _int, _int, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pv2_field_sample_bwd": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _int, _int, _int, _int, _int, _vp, _vp]),
    "pv2_sdf_mlp_param_count": (_i64, [_int, _int]),
    "pv2_sdf_mlp_fwd": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, C.c_float, _i64, _vp, _vp, _vp, _vp]),
    "pv2_sdf_mlp_bwd": (_int, [_vp, _vp, _vp, _int, _int, _int, _int, C.c_float, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _vp]),
    "pv2_ray_setup": (_int, [_vp, _vp, _vp, _int, _i64, _int, C.POINTER(C.c_float), C.c_float, _vp, _vp, _vp, _vp, _vp]),
    "pv2_ray_resample": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _i64, _int, _int, C.c_float, _int, C.c_float,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pv2_ray_composite_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, _i64, _int, _int, _vp, _vp, _vp,
                                      _vp, _vp]),
    "pv2_ray_composite_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, _i64, _int, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _vp, _vp]),
    "pv2_ray_loss_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, C.c_float, _vp, _vp]),
    "pv2_ray_loss_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pv2_bn_workspace_bytes": (_sz, [_i64, _int]),
    "pv2_bn_act_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _int, _i64, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pv2_bn_act_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pv2_bn_act_fwd_t": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _int, _i64, _int, _vp, _vp, _vp, _int, _vp, _sz,
                                 _vp]),
    "pv2_bn_act_bwd_t": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _i64, _int, _vp, _vp, _vp, _vp, _int, _int, _vp, _sz, _vp]),
    "pv2_densify_fwd": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _vp, _vp]),
    "pv2_densify_bwd": (_int, [_vp, _vp, _vp, _i64, _int, _i64, _vp, _vp]),
    "pv2_trilinear_fwd": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _int, _vp]),
    "pv2_trilinear_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _int, _vp]),
    "pv2_trilinear_bwd_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64,
                                     _int, _int, _int, _int, _vp]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load (once) and return the ctypes library; raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m ponderv2_b200.build` "
            "(ponderv2_b200 has no CPU or eager fallback)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError -> the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().pv2_error_string(code)
        raise RuntimeError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL) as a plain int (ctypes converts it for `c_void_p` parameters).
    The tensor must be CUDA + contiguous."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("ponderv2_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("ponderv2_b200 kernels need contiguous tensors")
    return t.data_ptr()


_torch_C = None


def _tc():
    global _torch_C
    if _torch_C is None:
        import torch
        _torch_C = torch._C
    return _torch_C


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device, as an int.  The wrappers ask ~400 times per step:
    `torch.cuda.current_stream()` builds a Stream object through several Python layers (13 us measured, 7 ms per
    step, profiles/r2k_host_profile.txt); the raw query is one C call."""
    tc = _tc()
    return tc._cuda_getCurrentRawStream(tc._cuda_getDevice())


_WS = {}


def workspace(nbytes: int, device):
    """Scratch of at least `nbytes` for a kernel launched on the CURRENT stream of `device`: one growing buffer per
    (device, stream), so back-to-back calls do not go through the allocator (stream order makes the reuse safe; a
    different stream gets a different buffer).  None for nbytes == 0."""
    if nbytes <= 0:
        return None
    import torch
    key = (device.index, stream_ptr())
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def i32x3(vals):
    arr = (C.c_int32 * 3)(*[int(v) for v in vals])
    return arr


DTYPE_CODE = {}


def dtype_code(dt) -> int:
    import torch
    if not DTYPE_CODE:
        DTYPE_CODE.update({torch.float32: 0, torch.bfloat16: 1, torch.float64: 2})
    try:
        return DTYPE_CODE[dt]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {dt}") from None


# ---------------------------------------------------------------------------------------------
# optional per-call timing (used by bench.py for the roofline of the dominant kernel)
# ---------------------------------------------------------------------------------------------
class _Profile:
    def __init__(self):
        self.enabled_for = None   # set of C-ABI names to time, or None
        self.records = []         # (name, start_event, end_event, algorithmic_bytes, flops)

    def start(self, names):
        self.enabled_for = set(names)
        self.records = []

    def stop(self):
        self.enabled_for = None

    def summary(self):
        """{name: dict(calls, ms, bytes, flops)} - call after torch.cuda.synchronize()."""
        out = {}
        for name, e0, e1, nbytes, flops in self.records:
            d = out.setdefault(name, dict(calls=0, ms=0.0, bytes=0, flops=0))
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += nbytes
            d["flops"] += flops
        return out


PROFILE = _Profile()


class _NullCtx:
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def on_device(dev):
    """`with on_device(t.device):` — torch.cuda.device(dev) only when `dev` is not already current (the context manager
    costs two driver calls per entry; the wrappers are entered ~1000 times per step)."""
    if dev.index is None or _tc()._cuda_getDevice() == dev.index:
        return _NULL
    import torch
    return torch.cuda.device(dev)


def timed(name, nbytes=0, flops=0):
    """with timed("pv2_x", bytes, flops): <C-ABI call>  — records CUDA events on the launching stream when enabled."""
    if PROFILE.enabled_for is None or name not in PROFILE.enabled_for:
        return _NULL
    return _Timed(name, nbytes, flops)


class _Timed:
    """CUDA-event pair around one C-ABI call (bench.py's instrumented loop)."""

    __slots__ = ("name", "nbytes", "flops", "e0")

    def __init__(self, name, nbytes=0, flops=0):
        self.name, self.nbytes, self.flops, self.e0 = name, nbytes, flops, None

    def __enter__(self):
        if PROFILE.enabled_for is not None and self.name in PROFILE.enabled_for:
            import torch
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            import torch
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.records.append((self.name, self.e0, e1, self.nbytes, self.flops))
        return False
