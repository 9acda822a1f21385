"""A/B builds of libjmb200.so for kernel experiments: jmb200.cu recompiled with extra -D flags, linked with the other
translation units of the regular build.  Usage: python scripts/build_variants.py tag1:-DX=1,-DY=2 tag2:-DZ=0 ...
Writes jmbayes_b200/libjmb200_<tag>.so (select with JMB_LIB=...)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmbayes_b200 import build  # noqa: E402

build.build_cuda()
nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]
others = [os.path.join(build.OBJ, u.replace(".cu", ".o")) for u in build.UNITS if u != "jmb200.cu"]


def one(spec):
    tag, _, defs = spec.partition(":")
    obj = os.path.join(build.OBJ, f"jmb200_{tag}.o")
    out = os.path.join(build.HERE, f"libjmb200_{tag}.so")
    subprocess.run([nvcc] + flags + [d for d in defs.split(",") if d] + ["-c", "-o", obj, os.path.join(build.CSRC, "jmb200.cu")], check=True, cwd=build.ROOT)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, obj] + others + ["-ldl"], check=True, cwd=build.ROOT)
    return out


with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)
