"""Build libconch_b200.so in-tree with nvcc for sm_100a (no torch headers, no pybind: a plain C-ABI library).

    python -m conch_b200.build [--force] [--verbose]

The built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import pathlib
import shutil
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# CB200_NVCC_EXTRA (e.g. -DCB200_TC_TRACE for benchmarks/tc_trace.py) builds a separate debug library next to the product
_DEBUG = bool(os.environ.get("CB200_NVCC_EXTRA", "").strip())
OBJ = PKG / ("build_dbg" if _DEBUG else "build")
LIB = PKG / ("libconch_b200_dbg.so" if _DEBUG else "libconch_b200.so")

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler",
    "-fPIC",
    "-Xcompiler",
    "-fvisibility=hidden",
    *os.environ.get("CB200_NVCC_EXTRA", "").split(),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and pathlib.Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libconch_b200.so cannot be built (there is no prebuilt or CPU fallback)")


def _signature(sources: list[pathlib.Path]) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(list(CSRC.glob("*.cuh")) + sources + [PKG.parent / "include" / "conch_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    sources = sorted(CSRC.glob("*.cu"))
    sig = _signature(sources)
    stamp = LIB.with_suffix(".so.sig")  # next to the library: the object directory does not travel to the GPU box
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == sig:
        return LIB
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)

    def compile_one(src: pathlib.Path) -> pathlib.Path:
        obj = OBJ / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(sig)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
