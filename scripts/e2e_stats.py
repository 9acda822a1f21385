"""Timing breakdown of the host-buffer path of jmb_mcmc_statistics (experiments)."""
import sys, time
import numpy as np, torch
from jmbayes_b200 import summaries
S, M = int(sys.argv[1]) if len(sys.argv) > 1 else 750000, 500
x = torch.randn((M, S), dtype=torch.float64).pin_memory()
xn = x.numpy()
w = np.full(M, 1.0 / M)
for rep in range(4):
    t0 = time.perf_counter()
    out, ms = summaries.series_statistics(xn, S, M, 1, S, w)
    print("wall %.1f ms, kernels %.1f ms" % (1e3 * (time.perf_counter() - t0), ms), flush=True)
for rep in range(2):
    t0 = time.perf_counter()
    out, ms = summaries.series_statistics(xn, S, M, 1, S, w, want=("postMeans",))
    print("one output: wall %.1f ms, kernels %.1f ms" % (1e3 * (time.perf_counter() - t0), ms), flush=True)
