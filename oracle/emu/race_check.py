"""Race check of the kernels' shared-memory / global-memory protocol: the emulated library (make_emu.py) built with
ThreadSanitizer.  TEST INFRASTRUCTURE ONLY.

In the emulation every CUDA thread is an OS thread and __syncthreads / __syncwarp / __shfl_*_sync are mutex + condition
variable barriers, so ThreadSanitizer sees exactly the happens-before edges the CUDA programming model guarantees: a missing
barrier between a shared-memory write of one lane and a read of another (which a B200 may tolerate by accident of warp
lock-step) is reported as a data race.  `python oracle/emu/race_check.py` prints the number of reports for a short chain of
every kernel family (MCMC step / finalize / Laplace, accuracy, posterior summaries, survfitJM); `--drop-syncwarp` is the negative control (one __syncwarp of the generic step kernel removed in a scratch copy of the
generated sources: the check must then report races).  (Removing the __syncwarp that guards the re-use of a stage of the
TMA-staged kernel is not a usable control: the emulated warp then deadlocks on the mbarrier phase, as the device would.)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_emu  # noqa: E402

DRIVER = r'''
import os, sys, ctypes as C
sys.path.insert(0, %(root)r)
os.environ["JMB_KERNEL"] = "generic"
import numpy as np
from jmbayes_b200 import api, accuracy
api._lib_cache = api._bind(C.CDLL(%(lib)r))
from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort
from oracle import oracle
for name in %(names)r:
    ms = name == "multistate"
    if not ms and %(extra)r:
        # the compile-time specialised kernel and the TMA-staged kernel (two-stage cp.async.bulk + mbarrier pipeline per warp)
        c = make_cohort(name, n=128)     # 2 CTAs x 8 warps x 8 subjects: 4 (16 lanes) or 8 (32 lanes) pipeline steps per warp
        ctl = dict(c["control"]); ctl.update(n_burnin=4, n_iter=2, n_block=2, n_thin=2)
        # None = the default: row-strided kernel; early = its staging variant (paired ring copies, group slots / stages handed
        # over between the lanes of a group: every variant's __syncwarp protocol is exercised)
        for mode, early, prefix in (("direct", None, "static:"), ("tma", None, "static-tma:"), (None, None, "static-rows"),
                                    (None, "0", "static-rows"), (None, "1", "static-rows"), (None, "2", "static-rows"), (None, "3", "static-rows")):
            if mode: os.environ["JMB_KERNEL"] = mode
            else: os.environ.pop("JMB_KERNEL", None)
            if early is None: os.environ.pop("JMB_ROWS_EARLY", None)
            else: os.environ["JMB_ROWS_EARLY"] = early
            with api.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"]) as fit:
                assert fit.layout_info()["kernel"].startswith(prefix), fit.layout_info()["kernel"]
                fit.set_draw(c["Data"])
                fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
        os.environ["JMB_KERNEL"] = "generic"; os.environ.pop("JMB_ROWS_EARLY", None)
    c = make_multistate_cohort(n=24) if ms else make_cohort(name, n=24)
    ctl = dict(c["control"]); ctl.update(n_burnin=6, n_iter=4, n_block=2, n_thin=2, adaptCov=True)
    pr = dict(c["priors"]); pr.update(shrink_gammas=True, shrink_alphas=True)
    with api.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"], multiState=ms) as fit:
        fit.set_xdesigns(c["Data"]); fit.set_draw_betas(c["Data"]["betas"], c["Data"]["sigmas"], c["Data"]["invD"])
        fit.lap_rwm_C(c["inits"], pr, c["scales"], c["Covs"], ctl)
        b = np.asarray(c["inits"]["b"])
        Wl = oracle.wlong(c["Data"], c["Covs"]["b"], c["interval_cens"], b, 0, multiState=ms)
        Wls = oracle.wlong(c["Data"], c["Covs"]["b"], c["interval_cens"], b, 1, multiState=ms)
        fit.set_wlong(Wl, Wls)
        fit.gradient_logPosterior(1.0, c["inits"]["Bs_gammas"], c["inits"]["gammas"], c["inits"]["alphas"], c["priors"], 2.0, 1.0, 0.01)
rng = np.random.default_rng(1)
n = 300
T = np.sort(rng.uniform(2, 9, n)); d = rng.choice([0.0, 1.0, 2.0, 3.0], size=n); p = rng.uniform(size=n)
accuracy.aucJM_pairs(T, d, 5.0, p, p, p); accuracy.rocJM(T, d, 5.0, p, p); accuracy.prederrJM_sum(p, p, p, p, 3 * n)
if %(extra)r:
    from jmbayes_b200 import survfit as sv, summaries
    from oracle import oracle_svft, oracle_stats
    X = np.cumsum(rng.normal(size=(40, 120)), axis=1) * 0.1 + rng.normal(size=(40, 120))
    summaries.series_statistics(np.ascontiguousarray(X), 40, 120, 120, 1, oracle_stats.stand(rng.normal(size=120)))
    c5 = make_cohort("config3", n=16)
    d5 = sv.newdata_designs(c5, L=3)
    means, draws = sv.posterior_draws(c5, M=6)
    with sv.SurvFit(d5) as h:
        g5 = h.modes(means)
        h.sample(draws, g5["modes"], g5["hessians"])
print("race-check driver done")
'''


def libtsan():
    r = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True)
    p = r.stdout.strip()
    return p if r.returncode == 0 and os.path.isabs(p) and os.path.exists(p) else None


def run(drop_syncwarp: bool = False, names=("config2", "config4", "multistate"), extra: bool = True):
    """Returns (number of ThreadSanitizer reports, report text)."""
    make_emu.build()
    work = tempfile.mkdtemp(prefix="jmb_tsan_")
    try:
        src = os.path.join(work, "src")
        shutil.copytree(make_emu.OUT, src)
        if drop_syncwarp:
            p = os.path.join(src, "jmb_kernels.cuh")
            s = open(p).read()
            a = "        for (int j = lane; j < QS; j += G) bo[j] = st[j];\n        __syncwarp(gmask);"
            assert s.count(a) == 1
            open(p, "w").write(s.replace(a, "        for (int j = lane; j < QS; j += G) bo[j] = st[j];"))
        lib = os.path.join(work, "libjmb200_emu_tsan.so")
        cmd = ["g++", "-O1", "-g", "-fsanitize=thread", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w", "-fpermissive", "-DJMB_EMU"] + os.environ.get("JMB_EMU_DEFS", "").split() + ["-I", os.path.join(HERE, "shim"),
               "-I", src, "-o", lib] + [os.path.join(src, f) for f in ("jmb200_emu.cpp", "jmb_laplace_emu.cpp", "jmb_accuracy_emu.cpp", "jmb_survfit_emu.cpp",
                                                                           "jmb_summaries_emu.cpp")] + \
              [os.path.join(HERE, "emu_runtime.cpp"), "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("TSan build failed:\n" + r.stderr[-4000:])
        env = dict(os.environ, LD_PRELOAD=libtsan(), TSAN_OPTIONS="report_signal_unsafe=0 halt_on_error=0 exitcode=0 history_size=2")
        r = subprocess.run([sys.executable, "-c", DRIVER % dict(root=ROOT, lib=lib, names=tuple(names), extra=bool(extra))], capture_output=True, text=True, env=env, cwd=work,
                           timeout=900)
        if "race-check driver done" not in r.stdout:
            raise RuntimeError("race-check driver failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
        return r.stderr.count("WARNING: ThreadSanitizer"), r.stderr
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    if libtsan() is None:
        sys.exit("libtsan.so not found")
    n, text = run(drop_syncwarp="--drop-syncwarp" in sys.argv)
    print("ThreadSanitizer reports:", n)
    if n:
        print(text[:6000])
