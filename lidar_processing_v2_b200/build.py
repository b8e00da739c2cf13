"""In-tree build of the CUDA library (sm_100a only, no other arch, no CPU fallback)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.path.join(HERE, "liblpl_b200.so")
SOURCES = ["capi.cu", "ring_dror.cu", "segment.cu", "cluster.cu", "hull.cu", "obb.cu", "ingest.cu", "split.cu", "knn.cu"]
HEADERS = ["common.cuh", "libm_exact.cuh", os.path.join("..", "..", "include", "lpl_b200.h")]

# -fmad=false / -ffp-contract=off: the reference binary (x86-64 baseline, no FMA) evaluates
# every float expression without contraction; parity is bit-exact only if we do the same.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-shared", "-cudart", "static", "--threads", "4",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    so_m = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(p) and os.path.getmtime(p) > so_m for p in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into lidar_processing_v2_b200/liblpl_b200.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return SO_PATH
    extra = os.environ.get("LPL_NVCC_EXTRA", "").split()  # e.g. -DLPL_HULL_TRACE (diagnostics only)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", SO_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose and (res.stdout or res.stderr):
        print(res.stdout, res.stderr)
    return SO_PATH


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
