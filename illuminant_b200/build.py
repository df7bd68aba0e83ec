"""Builds libilluminant_b200.so (sm_100a only) in-tree with nvcc.

`python -m illuminant_b200.build` or `illuminant_b200.build.build()`; invoked by `__graft_entry__.build()`.
The product has no CPU fallback: if nvcc is missing the build fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = Path(os.environ["ILB_OUT"]) if os.environ.get("ILB_OUT") else PKG / "libilluminant_b200.so"  # ILB_OUT: variant builds
SOURCES = ["api.cu", "lighting.cu", "particles.cu", "dfgen.cu", "planes.cu", "resolve.cu", "raster.cu"]
HEADERS = ["ilb_device.cuh", "ilb_internal.h", "ilb_shapes.cuh", "ilb_bezier.cuh", "../../include/illuminant_b200.h", "../../include/ilb_detmath.h"]

# FMA contraction on, 2-ulp division / sqrt (MUFU based), denormals flushed; sin/cos/pow/atan2/acos stay the accurate
# library versions (no -use_fast_math).  See DESIGN.md "Numerics".  ILB_EXACT=1 builds the bit-faithful variant
# (-fmad=false, IEEE div/sqrt) used to separate arithmetic differences from logic differences when debugging parity.
EXACT = os.environ.get("ILB_EXACT") == "1"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--threads", "0",   # the translation units compile in parallel
    *(["-fmad=false"] if EXACT else ["-fmad=true", "-prec-div=false", "-prec-sqrt=false", "-ftz=true"]),
    *[f"-D{d}" for d in os.environ.get("ILB_DEFINES", "").split() if d],
    "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-O2",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libilluminant_b200.so cannot be built (there is no CPU fallback)")


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS] + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++", "-o", str(LIB), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr, file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
