"""In-tree build of the CUDA library (sm_100a only)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libjmb200.so")
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("jmb200.cu", "jmb_kernels.cuh", "jmb_device.cuh")]
HEADERS = [os.path.join(ROOT, "include", "jmb200.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -> jmbayes_b200/libjmb200.so."""
    if not force and not _stale(LIB, SOURCES + HEADERS):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SOURCES[0], "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=ROOT)
    return LIB
