"""Build libcsb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

`python -m commonscenes_b200.build` compiles every .cu under csrc/ into
commonscenes_b200/libcsb200.so.  nvcc cross-compiles without a GPU, so this runs in the build
container; the .so travels to the GPU box with the repo snapshot (it is git-ignored, not
gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
BUILD = ROOT / "_build"
LIB = ROOT / "libcsb200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                      list((ROOT.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    h.update(" ".join(ARCH_FLAGS + COMMON).encode())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".sha")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [NVCC, *ARCH_FLAGS, *COMMON, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (BUILD / (src.stem + ".ptxas.log")).write_text(res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {src.name}")
    if verbose:
        for line in res.stderr.splitlines():
            if "registers" in line or "spill" in line and "0 bytes spill" not in line:
                print(f"[{src.name}] {line.strip()}")
    stamp.write_text(dig)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    if force:
        for f in BUILD.glob("*.sha"):
            f.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, *ARCH_FLAGS, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    path = build(verbose=True, force="--force" in sys.argv)
    print(path)
