"""In-tree build of the CUDA library (sm_100a only)."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libjmb200.so")
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
# translation units and what each one includes
UNITS = {
    "jmb200.cu": ["jmb_kernels.cuh", "jmb_device.cuh", "jmb_model.h", "jmb_step_static.cuh", "jmb_step_rows.cuh", "jmb_host.h"],
    "jmb_survfit.cu": ["jmb_survfit.cuh", "jmb_device.cuh", "jmb_model.h", "jmb_host.h"],
    "jmb_laplace.cu": ["jmb_host.h"],
    "jmb_summaries.cu": ["jmb_host.h"],
    "jmb_accuracy.cu": ["jmb_host.h"],
    "jmb_designs.cu": ["jmb_host.h"],
}
HEADERS = [os.path.join(ROOT, "include", f) for f in ("jmb200.h", "jmb200_survfit.h", "jmb200_summaries.h", "jmb200_accuracy.h", "jmb200_designs.h")]
SOURCES = [os.path.join(CSRC, f) for f in UNITS]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -> jmbayes_b200/libjmb200.so."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
             "-Xcompiler", "-fPIC"]
    if verbose:
        flags.insert(0, "-Xptxas=-v")
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for unit, incs in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        deps = [src] + [os.path.join(CSRC, i) for i in incs] + HEADERS
        if force or _stale(obj, deps):
            jobs.append([nvcc] + flags + ["-c", "-o", obj, src])
    if jobs:
        with ThreadPoolExecutor(len(jobs)) as ex:
            for r in ex.map(lambda c: subprocess.run(c, cwd=ROOT, capture_output=not verbose, text=True), jobs):
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed:\n" + (r.stderr or "") + (r.stdout or ""))
                if verbose and r.stderr:
                    print(r.stderr)
    objs = [os.path.join(OBJ, u.replace(".cu", ".o")) for u in UNITS]
    if jobs or _stale(LIB, objs):
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-ldl"],
                       check=True, cwd=ROOT)
    return LIB
