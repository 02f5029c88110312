import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if needed) and load libpv2_b200.so; GPU tests call the product through it."""
    from ponderv2_b200 import _lib, build
    build.build_cuda()
    return _lib.load()


def record(name: str, **values) -> None:
    """Append measured parity numbers to gpurun_out/parity_report.jsonl (copied to profiles/ per round): the tests
    assert tolerances, this keeps the actual errors on record."""
    import json
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        with open(out / "parity_report.jsonl", "a") as f:
            f.write(json.dumps({"test": name, **{k: (float(v) if isinstance(v, (int, float)) else v)
                                                 for k, v in values.items()}}) + "\n")
    except OSError:
        pass
