"""Builds oracle/emu/_build/libjmb200_emu.so: the product's own jmb200.cu + kernels compiled by g++ against the stand-in
CUDA runtime of oracle/emu/shim (kernels run on CPU threads).  TEST INFRASTRUCTURE ONLY - see shim/cuda_runtime.h.

The sources are used where they lie; three textual substitutions make them C++:
  1. `kernel<<<grid, block, smem, stream>>>(args)`  ->  `emu::launch(grid, block, smem, stream, [&] { kernel(args); })`
  2. `extern __shared__ T name[];` -> `T* name = (T*)emu::dyn_smem();`, `__shared__` -> `static`   (one CTA runs at a time)
  3. the `asm volatile("griddepcontrol...")` / `fence.mbarrier_init` statements -> empty statements; the four PTX helpers of
     the TMA-staged kernel (mbarrier.init / arrive.expect_tx / try_wait, cp.async.bulk) -> shim/emu_mbarrier.h;
     `__noinline__` -> `__attribute__((noinline))`
"""
from __future__ import annotations

import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "jmbayes_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libjmb200_emu.so")
SOURCES = ["jmb200.cu", "jmb_kernels.cuh", "jmb_device.cuh", "jmb_model.h", "jmb_host.h", "jmb_laplace.cu", "jmb_accuracy.cu",
           "jmb_survfit.cu", "jmb_survfit.cuh", "jmb_summaries.cu", "jmb_step_static.cuh", "jmb_step_rows.cuh"]
CALLEE = re.compile(r"[A-Za-z_][A-Za-z_0-9]*(?:->[A-Za-z_][A-Za-z_0-9]*)*(?:<[^<>;(){}]*>)?$")


def _match(text, i, open_c, close_c):
    depth = 0
    for j in range(i, len(text)):
        if text[j] == open_c:
            depth += 1
        elif text[j] == close_c:
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced")


def rewrite_launches(text: str) -> str:
    out, pos = [], 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            return "".join(out)
        m = CALLEE.search(text[pos:k])
        if not m:
            raise ValueError("cannot find the kernel name before <<< at " + text[max(0, k - 80):k])
        callee_start = pos + m.start()
        e = text.index(">>>", k)
        cfg = text[k + 3:e]
        a0 = e + 3
        assert text[a0] == "(", text[a0:a0 + 20]
        a1 = _match(text, a0, "(", ")")
        parts, depth, cur = [], 0, ""
        for ch in cfg:
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur.strip()); cur = ""
            else:
                cur += ch
        parts.append(cur.strip())
        while len(parts) < 4:
            parts.append("0" if len(parts) == 2 else "nullptr")
        out.append(text[pos:callee_start])
        out.append("emu::launch(%s, %s, %s, %s, [&] { %s%s; })" % (parts[0], parts[1], parts[2], parts[3], text[callee_start:k], text[a0:a1 + 1]))
        pos = a1 + 1


def transform(text: str) -> str:
    text = rewrite_launches(text)
    text = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::dyn_smem();", text)
    text = text.replace("__shared__", "static").replace("__noinline__", "__attribute__((noinline))")
    text = re.sub(r'asm volatile\("(?:griddepcontrol|fence\.mbarrier_init|cp\.async\.bulk\.prefetch)[^;]*;"[^;]*;', ";", text)
    # the PTX helpers of the TMA-staged kernel (mbarrier + cp.async.bulk) -> their emulation (shim/emu_mbarrier.h)
    i = text.find("__device__ __forceinline__ uint32_t smem_u32")
    if i >= 0:
        j = text.index("struct TmaArgs {", i)
        text = text[:i] + '#include "emu_mbarrier.h"\n\n' + text[j:]
    text = text.replace('"../../include/', '"' + os.path.join(ROOT, "include") + "/")
    return text


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "shim", f) for f in os.listdir(os.path.join(HERE, "shim"))] + \
           [os.path.join(HERE, "emu_runtime.cpp"), os.path.abspath(__file__), os.path.join(ROOT, "include", "jmb200.h"),
            os.path.join(ROOT, "include", "jmb200_accuracy.h"),
            os.path.join(ROOT, "include", "jmb200_survfit.h"), os.path.join(ROOT, "include", "jmb200_summaries.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    for s in SOURCES:
        dst = os.path.join(OUT, s.replace(".cu", "_emu.cpp") if s.endswith(".cu") else s)
        with open(os.path.join(CSRC, s)) as f:
            txt = transform(f.read())
        with open(dst, "w") as f:
            f.write(txt)
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w", "-fpermissive", "-fmax-errors=15", "-DJMB_EMU"] + os.environ.get("JMB_EMU_DEFS", "").split() + \
          ["-I", os.path.join(HERE, "shim"), "-I", OUT,
           "-o", LIB, os.path.join(OUT, "jmb200_emu.cpp"), os.path.join(OUT, "jmb_laplace_emu.cpp"), os.path.join(OUT, "jmb_accuracy_emu.cpp"),
           os.path.join(OUT, "jmb_survfit_emu.cpp"), os.path.join(OUT, "jmb_summaries_emu.cpp"),
           os.path.join(HERE, "emu_runtime.cpp"), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + r.stderr[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
