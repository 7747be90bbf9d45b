"""TEST INFRASTRUCTURE.  Builds the checkers that are compiled code:

* ``oracle/_build/liboracle_points.so`` -- gcc build of ``oracle/points.c`` (the CPU restatement of the reference's
  point-cloud kernels);
* ``oracle/_ref/libref_points.so`` -- the reference's OWN CUDA kernels for that path
  (``scripts/pytorch_structural_losses/src/{approxmatch,nndistance}.cu``: self-contained, no ATen), compiled with nvcc for
  sm_100a from the sources where they lie under /root/reference, plus the forwarding shim ``oracle/ref_points/shim.cu``.
  Only possible where /root/reference exists (the build container); the .so travels to the GPU box with the snapshot
  (git-ignored, not gpurun-ignored).  ``extension/chamfer.cu`` holds the same NmDistance kernels behind ATen tensors and is
  not built (it needs the torch C++ headers and the long-removed ``Tensor::data<T>()``).

Nothing under commonscenes_b200/ imports or links any of this.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path(os.environ.get("CS_REFERENCE_ROOT", "/root/reference")) / "scripts" / "pytorch_structural_losses" / "src"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ORACLE_LIB = HERE / "_build" / "liboracle_points.so"
REF_LIB = HERE / "_ref" / "libref_points.so"


def _stale(target: Path, sources) -> bool:
    return not target.exists() or any(Path(s).stat().st_mtime > target.stat().st_mtime for s in sources)


def build_oracle_c(verbose: bool = False) -> Path:
    src = HERE / "points.c"
    if _stale(ORACLE_LIB, [src]):
        ORACLE_LIB.parent.mkdir(exist_ok=True)
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", str(src), "-o", str(ORACLE_LIB), "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return ORACLE_LIB


def build_reference_points(verbose: bool = False) -> Path | None:
    """None when the reference sources are not on this machine (the GPU box): the prebuilt library is used if present."""
    srcs = [REF_SRC / "approxmatch.cu", REF_SRC / "nndistance.cu"]
    if not all(s.exists() for s in srcs):
        return REF_LIB if REF_LIB.exists() else None
    shim = HERE / "ref_points" / "shim.cu"
    if _stale(REF_LIB, [*srcs, shim]):
        REF_LIB.parent.mkdir(exist_ok=True)
        cmd = [NVCC, "-O3", "-DNDEBUG", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
               f"-I{REF_SRC}", *map(str, srcs), str(shim), "-o", str(REF_LIB), "-lcudart"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return REF_LIB


if __name__ == "__main__":
    print(build_oracle_c(verbose=True))
    print(build_reference_points(verbose=True))
