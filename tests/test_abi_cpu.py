"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/pv2_b200.h declares; the Python shims fail loudly without a GPU."""
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "pv2_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pv2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(cuda_lib):
    from ponderv2_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(cuda_lib, n), f"{n} declared in include/pv2_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert cuda_lib.pv2_version() >= 100
    assert cuda_lib.pv2_error_string(-1).decode().startswith("pv2")
    assert cuda_lib.pv2_rulebook_workspace_bytes(1000) > 1000 * 12


def test_no_cpu_fallback():
    import ponderv2_b200.spconv.pytorch as spconv
    from ponderv2_b200.smooth_sampler import SmoothSampler
    with pytest.raises(RuntimeError):
        spconv.SparseConvTensor(torch.zeros(2, 4), torch.zeros(2, 4, dtype=torch.int32), [4, 4, 4], 1)
    with pytest.raises(RuntimeError):
        SmoothSampler.apply(torch.rand(1, 2, 3, 3, 3), torch.rand(1, 1, 1, 4, 3), "zeros", True, False)


def test_product_never_imports_oracle():
    for p in (ROOT / "ponderv2_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_backbone_state_dict_contract_cpu():
    import json
    from ponderv2_b200.backbone import SpUNetBase
    want = json.loads((ROOT / "tests" / "golden" / "spunet_v1m1_state.json").read_text())
    m = SpUNetBase(in_channels=6, num_classes=0)
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == want["state"]
    assert sum(p.numel() for n, p in m.named_parameters() if p.dim() == 5) == want["conv_params"] == 39138752
