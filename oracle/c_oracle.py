"""ctypes binding of oracle/spconv_c.c (TEST INFRASTRUCTURE, see oracle/__init__.py): a second, plain-C restatement of
the rulebook and sparse-convolution semantics, cross-checked against oracle/spconv_oracle.py in tests/test_oracle_cpu.py.
Built by ponderv2_b200.build.build_oracle_c() into oracle/lib/liboracle.so."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_LIB = Path(__file__).resolve().parent / "lib" / "liboracle.so"
_lib = None


def load():
    global _lib
    if _lib is None:
        if not _LIB.exists():
            raise RuntimeError(f"{_LIB} missing: run `python -m ponderv2_b200.build` (builds the oracle's C restatement)")
        _lib = C.CDLL(str(_LIB))
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        _lib.oc_subm_rulebook.argtypes = [vp, i64, vp, i32, vp]
        _lib.oc_down_rulebook.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        _lib.oc_sparse_conv_f64.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, i32]
        for f in (_lib.oc_subm_rulebook, _lib.oc_down_rulebook, _lib.oc_sparse_conv_f64):
            f.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def subm_rulebook(coords: np.ndarray, shape, ksize: int) -> np.ndarray:
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    n = coords.shape[0]
    nbr = np.empty((ksize ** 3, n), dtype=np.int32)
    sh = np.asarray(shape, dtype=np.int32)
    assert load().oc_subm_rulebook(_p(coords), n, _p(sh), ksize, _p(nbr)) == 0
    return nbr


def down_rulebook(coords: np.ndarray, shape):
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    n = coords.shape[0]
    out = np.empty((max(n, 1), 4), dtype=np.int32)
    in2out = np.empty(max(n, 1), dtype=np.int32)
    koff = np.empty(max(n, 1), dtype=np.int32)
    n_out = np.zeros(1, dtype=np.int64)
    sh = np.asarray(shape, dtype=np.int32)
    assert load().oc_down_rulebook(_p(coords), n, _p(sh), _p(out), _p(in2out), _p(koff), _p(n_out)) == 0
    m = int(n_out[0])
    return out[:m], in2out[:n], koff[:n], [(int(s) - 2) // 2 + 1 for s in shape]


def sparse_conv(x: np.ndarray, w: np.ndarray, bias, nbr: np.ndarray) -> np.ndarray:
    """x [n_in, cin] f64, w [cout, K, cin] f64, nbr [K, n_out] int32 -> y [n_out, cout] f64."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    nbr = np.ascontiguousarray(nbr, dtype=np.int32)
    cout, kvol, cin = w.shape
    y = np.empty((nbr.shape[1], cout), dtype=np.float64)
    b = np.ascontiguousarray(bias, dtype=np.float64) if bias is not None else None
    assert load().oc_sparse_conv_f64(_p(x), _p(w), _p(b), _p(nbr), _p(y), nbr.shape[1], cin, cout, kvol) == 0
    return y
