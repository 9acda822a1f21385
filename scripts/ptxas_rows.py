"""Registers / spills of every jmb_step_rows instantiation: python scripts/ptxas_rows.py [extra -D flags]"""
import re, subprocess, sys
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
       "-Xcompiler", "-fPIC", "-Xptxas=-v", "-c", "-o", "/tmp/jmb200_v.o", "jmbayes_b200/csrc/jmb200.cu"] + sys.argv[1:]
t = subprocess.run(cmd, capture_output=True, text=True).stderr
for m in re.finditer(r"Compiling entry function '(\S+)' for[^\n]*\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n[^\n]*Used (\d+) registers", t):
    k = re.search(r"jmb_step_rowsINS_(\d+)(\w+?)ELi(\d+)ELi(\d+)EEE", m.group(1))
    if k:
        name = k.group(2)[:int(k.group(1))]
        print(f"{name:22s} G={k.group(3):>2s} early={k.group(4)}  regs {m.group(5)}  stack {m.group(2):>3s}  spill st/ld {m.group(3):>3s}/{m.group(4):>3s}")
