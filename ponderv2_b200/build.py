"""In-tree nvcc build of libpv2_b200.so (sm_100a only) and of the oracle's C restatement.

`python -m ponderv2_b200.build` or `__graft_entry__.build()`.  Objects go to build/, the shared
library to ponderv2_b200/lib/ so that it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "ponderv2_b200" / "csrc"
LIBDIR = ROOT / "ponderv2_b200" / "lib"
OBJDIR = ROOT / "build" / "obj"
LIBNAME = "libpv2_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]
if os.environ.get("PV2_MBAR_DEBUG") == "1":   # development: non-fatal barrier watchdog with a wait log (csrc/umma.cuh)
    NVCC_FLAGS.append("-DPV2_MBAR_DEBUG")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; ponderv2_b200 cannot be built")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_cuda(verbose: bool = False, force: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    LIBDIR.mkdir(parents=True, exist_ok=True)
    OBJDIR.mkdir(parents=True, exist_ok=True)
    lib = LIBDIR / LIBNAME
    stamp = LIBDIR / (LIBNAME + ".sha256")
    hdr_digest = _digest(headers)
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        tag = OBJDIR / (src.stem + ".sha256")
        want = _digest([src]) + hdr_digest
        if not force and obj.exists() and tag.exists() and tag.read_text() == want:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        tag.write_text(want)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    want_all = _digest(sources) + hdr_digest
    if force or not lib.exists() or not stamp.exists() or stamp.read_text() != want_all:
        cmd = [nvcc, "-shared", "-o", str(lib), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(want_all)
    return lib


def build_oracle_c() -> Path | None:
    """Compile oracle/*.c (CPU restatement used only by tests and bench's cpu_baseline)."""
    odir = ROOT / "oracle"
    srcs = sorted(odir.glob("*.c"))
    if not srcs:
        return None
    out = odir / "lib"
    out.mkdir(exist_ok=True)
    lib = out / "liboracle.so"
    stamp = out / "liboracle.sha256"
    want = _digest(srcs)
    if lib.exists() and stamp.exists() and stamp.read_text() == want:
        return lib
    # no -march=native: the library is built in one container and may be loaded on another host
    cmd = ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", str(lib), *map(str, srcs), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stderr}")
    stamp.write_text(want)
    return lib


def main() -> None:
    verbose = "-v" in sys.argv
    force = "-f" in sys.argv
    lib = build_cuda(verbose=verbose, force=force)
    print(f"built {lib}")
    o = build_oracle_c()
    if o:
        print(f"built {o}")


if __name__ == "__main__":
    main()
