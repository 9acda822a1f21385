"""CPU suite of the survfitJM path (BASELINE config 5): the restated R loop (oracle/jmb_oracle_svft.c) against
independent NumPy / SciPy evaluations and analytic properties, the host-side design builder, and the C ABI of
include/jmb200_survfit.h (symbols, struct layouts, loud failure without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from jmbayes_b200 import survfit as sv
from jmbayes_b200.cohorts import make_cohort

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def osv(oracle):
    from oracle import oracle_svft
    oracle_svft.lib()
    return oracle_svft


@pytest.fixture(scope="module")
def case3():
    c = make_cohort("config3", n=48)
    d = sv.newdata_designs(c, L=6)
    means, draws = sv.posterior_draws(c, M=40)
    return c, d, means, draws


def _np_log_post(d, par, m, i, b):
    """Independent NumPy statement of log_post_RE_svft (src/survfitJM_mvJMbayes.cpp:140-166) for subject i."""
    O, K = len(d["y_long"]), d["K"]
    betas = [np.asarray(x)[m] for x in par["betas"]]
    invD = np.asarray(par["inv_D"])[m]
    Bs, gam, alp = np.asarray(par["Bs_gammas"])[m], np.asarray(par["gammas"])[m], np.asarray(par["alphas"])[m]
    lp = 0.0
    for k in range(O):
        rows = np.nonzero(np.asarray(d["idL"][k]) == i)[0]
        if rows.size == 0:
            continue
        eta = np.asarray(d["X_long"][k])[rows] @ betas[k] + np.asarray(d["Z_long"][k])[rows] @ b[np.asarray(d["RE_inds"][k])]
        y = np.asarray(d["y_long"][k])[rows]
        fam = d["fams"][k]
        if fam == "gaussian":
            lp += np.sum(-0.5 * ((y - eta) / np.asarray(par["sigmas"][k])[m]) ** 2)
        elif fam == "binomial":
            pr = np.exp(eta) / (1 + np.exp(eta))
            lp += np.sum(y * np.log(pr) + (1 - y) * np.log(1 - pr))
        else:
            lp += np.sum(y * eta - np.exp(eta))
    lp += -0.5 * b @ invD @ b
    r = slice(i * K, (i + 1) * K)
    Wl = np.zeros((K, len(alp)))
    for t in range(len(d["XXs"])):
        k = d["outcome"][t]
        v = np.asarray(d["XXs"][t])[r] @ betas[k][np.asarray(d["indFixed"][t])] + np.asarray(d["ZZs"][t])[r] @ b[np.asarray(d["RE_inds2"][t])]
        Wl[:, np.asarray(d["col_inds"][t])] = np.asarray(d["Us"][t])[r] * v[:, None]
    lin = np.asarray(d["W1s"])[r] @ Bs + np.asarray(d["W2s"])[r] @ gam + Wl @ alp
    return lp - np.sum(np.asarray(d["Pw"])[r] * np.exp(lin))


def test_log_post_matches_numpy_and_the_pinned_restatement(osv, oracle, case3):
    c, d, means, draws = case3
    rng = np.random.default_rng(3)
    for i in (0, 5, 17, 47):
        for par, m in ((means, 0), (draws, 7), (draws, 39)):
            b = 0.3 * rng.normal(size=6)
            got = osv.log_post(d, par, m, i, b)
            ref = _np_log_post(d, par, m, i, b)
            assert abs(got - ref) <= 1e-11 * max(1.0, abs(ref))
    # tie to orc_eval_logpost_re, which tests/test_reference_golden.py pins to the reference's own source:
    # the same new subjects expressed as a lap_rwm_C-style Data list (Xbetas materialised on the host)
    n, K = d["n"], d["K"]
    O = len(d["y_long"])
    betas = [np.asarray(x)[0] for x in means["betas"]]
    cnt = [np.bincount(d["idL"][k], minlength=n) for k in range(O)]
    D = dict(c["Data"])
    D.update(y=d["y_long"], Z=d["Z_long"], idL=d["idL"], idL2=[(np.cumsum(x) - 1).astype(np.int32) for x in cnt],
             Xbetas=[np.asarray(d["X_long"][k]) @ betas[k] for k in range(O)],
             W1s=d["W1s"], W2s=d["W2s"], Pw=d["Pw"], ZZs=d["ZZs"], Us=d["Us"],
             XXsbetas=[np.asarray(d["XXs"][t]) @ betas[d["outcome"][t]][d["indFixed"][t]] for t in range(len(d["XXs"]))],
             sigmas=[float(means["sigmas"][k][0]) if means["sigmas"][k] is not None else float("nan") for k in range(O)],
             invD=np.asfortranarray(means["inv_D"][0]))
    b = 0.2 * rng.normal(size=(n, 6))
    ev = oracle.eval_logpost_re(D, c["Covs"]["b"], False, b, means["Bs_gammas"][0], means["gammas"][0], means["alphas"][0])
    for i in range(0, n, 7):
        got = osv.log_post(d, means, 0, i, b[i])
        ref = (ev["log_pyb"][i] + (0.0 - ev["H"][i])) + ev["log_pb"][i]
        assert abs(got - ref) <= 1e-10 * max(1.0, abs(ref))


def test_surv_pred_is_minus_the_interval_hazard(osv, case3):
    c, d, means, draws = case3
    b = np.array([0.1, -0.05, 0.2, 0.0, -0.1, 0.05])
    K, L = d["K"], d["L"]
    for i, l in ((0, 0), (3, 5), (40, 2)):
        got = osv.surv_pred(d, draws, 11, i, l, b)
        r = slice((i * L + l) * K, (i * L + l + 1) * K)
        betas = [np.asarray(x)[11] for x in draws["betas"]]
        Wl = np.zeros((K, 6))
        for t in range(6):
            k = d["outcome"][t]
            v = np.asarray(d["XXs_h"][t])[r] @ betas[k][d["indFixed"][t]] + np.asarray(d["ZZs_h"][t])[r] @ b[d["RE_inds2"][t]]
            Wl[:, d["col_inds"][t]] = np.asarray(d["Us_h"][t])[r] * v[:, None]
        lin = np.asarray(d["W1s_h"])[r] @ draws["Bs_gammas"][11] + np.asarray(d["W2s_h"])[r] @ draws["gammas"][11] + Wl @ draws["alphas"][11]
        ref = -np.sum(np.asarray(d["Pw_h"])[r] * np.exp(lin))
        assert got < 0 and abs(got - ref) <= 1e-12 * abs(ref)


def test_dense_helpers_against_numpy(osv):
    rng = np.random.default_rng(1)
    for n in (1, 2, 6, 8):
        A = rng.normal(size=(n, n)); S = A @ A.T + 0.1 * np.eye(n)
        inv, rc = osv.inverse(S)
        assert rc == 0 and np.allclose(inv, np.linalg.inv(S), rtol=1e-10, atol=1e-12)
        val, vec = osv.eigen_sym(S)
        assert np.allclose(val, np.sort(np.linalg.eigvalsh(S))[::-1], rtol=1e-12)
        assert np.allclose(vec @ np.diag(val) @ vec.T, S, rtol=1e-11, atol=1e-12)
        assert np.allclose(vec.T @ vec, np.eye(n), atol=1e-12)
    _, rc = osv.inverse(np.zeros((3, 3)))
    assert rc == 1


def test_summary_is_r_type7(osv):
    rng = np.random.default_rng(2)
    for M in (1, 2, 7, 200, 201):
        x = rng.uniform(size=M)
        s = osv.summary(x)
        assert np.isclose(s[0], x.mean(), rtol=1e-14) and np.isclose(s[1], np.median(x), rtol=1e-14)
        assert np.allclose(s[2:], np.quantile(x, [0.025, 0.975]), rtol=1e-13)     # numpy's default is R's type 7
    x = rng.uniform(size=50); x[[3, 10]] = np.nan                                   # na.rm = TRUE
    s = osv.summary(x)
    ok = x[~np.isnan(x)]
    assert np.isclose(s[0], ok.mean()) and np.isclose(s[1], np.median(ok)) and np.allclose(s[2:], np.quantile(ok, [0.025, 0.975]))
    assert np.all(np.isnan(osv.summary(np.full(5, np.nan))))


def test_mode_search_finds_the_posterior_mode(osv, case3):
    """vmmin restatement: convergence code 0, the optimiser's own gradient is ~0 at the mode, SciPy's BFGS on
    the same objective lands on the same point, and the optim Hessian matches a finite-difference Hessian."""
    from scipy.optimize import minimize
    c, d, means, draws = case3
    r = osv.modes(d, means)
    assert np.all(r["convergence"] == 0)
    assert np.all(r["counts"][1] >= 5) and np.all(r["counts"][1] < 80)             # BFGS iterations (gradient count)
    for i in (0, 9, 31):
        f = lambda b: -osv.log_post(d, means, 0, i, b)
        res = minimize(f, np.zeros(6), method="BFGS", options=dict(gtol=1e-7))
        assert np.allclose(res.x, r["modes"][i], atol=3e-3)        # optim stops on reltol = 1.5e-8 of f, not on the gradient
        assert f(r["modes"][i]) <= res.fun + 1e-6
        H = np.zeros((6, 6)); h = 1e-4; b0 = r["modes"][i]
        for a in range(6):
            for bb in range(6):
                ea, eb = h * np.eye(6)[a], h * np.eye(6)[bb]
                H[a, bb] = (f(b0 + ea + eb) - f(b0 + ea - eb) - f(b0 - ea + eb) + f(b0 - ea - eb)) / (4 * h * h)
        assert np.allclose(r["hessians"][:, :, i], H, rtol=2e-3, atol=2e-3)
        assert np.all(np.linalg.eigvalsh(r["hessians"][:, :, i]) > 0)


def test_sampler_properties(osv, case3):
    c, d, means, draws = case3
    r = osv.modes(d, means)
    s = osv.sample(d, draws, r["modes"], r["hessians"])
    full = s["full_results"]                                  # L x M x n
    assert full.shape == (6, 40, 48) and np.all(np.isfinite(full))
    assert np.all(full > 0) and np.all(full <= 1.0) and np.all(np.diff(full, axis=0) <= 0)   # exp(cumsum(logS))
    assert 0.2 < s["accept_rate"].mean() < 0.9
    for i in (0, 20):
        for l in (0, 5):
            x = full[l, :, i]
            assert np.allclose(s["summaries"][l, :, i], [x.mean(), np.median(x), *np.quantile(x, [0.025, 0.975])], rtol=1e-12)
    s2 = osv.sample(d, draws, r["modes"], r["hessians"])
    assert np.array_equal(s2["full_results"], full)           # counter RNG: deterministic
    s3 = osv.sample(d, draws, r["modes"], r["hessians"], control=dict(seed=2))
    assert not np.array_equal(s3["full_results"], full)
    sl = osv.sample(d, draws, r["modes"], r["hessians"], control=dict(log=True))
    assert np.allclose(np.exp(np.cumsum(sl["full_results"], axis=0)), full, rtol=1e-12)
    # a non-PD Hessian is reported, not propagated
    Hbad = r["hessians"].copy(); Hbad[:, :, 4] = -np.eye(6)
    sb = osv.sample(d, draws, r["modes"], Hbad)
    assert sb["convergence"][4] == 20 and np.all(np.isnan(sb["summaries"][:, :, 4])) and np.all(np.isfinite(sb["summaries"][:, :, 5]))
    # threads: subjects are independent
    a = osv.survfitJM(d, means, draws, threads=1)
    b = osv.survfitJM(d, means, draws, threads=3)
    assert np.array_equal(a["summaries"], b["summaries"])


def test_newdata_designs_follow_the_reference_construction():
    c = make_cohort("config3", n=30)
    d = sv.newdata_designs(c, L=4, horizon_step=0.5)
    K, n = d["K"], d["n"]
    lt = d["last_time"]
    assert np.allclose(d["Pw"].reshape(n, K).sum(axis=1), lt)                       # sum(P wk) = last.time (wk sums to 2)
    assert np.allclose(d["Pw_h"].reshape(n, 4, K).sum(axis=2), 0.5)
    st = np.asarray(d["ZZs"][0])[:, 1].reshape(n, K)
    assert np.all(st > 0) and np.all(st < lt[:, None])
    sth = np.asarray(d["ZZs_h"][0])[:, 1].reshape(n, 4, K)
    assert np.all(sth > lt[:, None, None]) and np.all(sth < (lt + 2.0)[:, None, None])
    for k in range(3):
        t = np.asarray(d["X_long"][k])[:, 1]
        assert np.all(t <= lt[d["idL"][k]] + 1e-12)
    rs = np.asarray(d["W1s"]).sum(axis=1)                                           # partition of unity inside the knots,
    assert np.all(np.isclose(rs, 1.0) | (rs == 0.0)) and np.mean(rs == 0.0) < 0.05  # zero rows outside (outer.ok = TRUE)


# ---- the C ABI ------------------------------------------------------------------------------------------
def test_cabi_exports_the_survfit_symbols_and_struct_layouts():
    import subprocess
    import tempfile
    from jmbayes_b200 import build
    lib = build.build_cuda()
    hdr = open(os.path.join(ROOT, "include", "jmb200_survfit.h")).read()
    names = sorted(set(re.findall(r"\b(jmb_s[a-zA-Z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 6
    L = C.CDLL(lib)
    for nm in names:
        assert hasattr(L, nm), nm
    src = ('#include <stdio.h>\n#include "jmb200_survfit.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(jmb_svft_data), '
           'sizeof(jmb_svft_params), sizeof(jmb_svft_control), sizeof(jmb_svft_out));return 0;}')
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")], check=True)
        out = subprocess.run([os.path.join(td, "s")], capture_output=True, text=True, check=True).stdout.split()
    got = [C.sizeof(x) for x in (sv.SvftData, sv.SvftParams, sv.SvftControl, sv.SvftOut)]
    assert got == [int(x) for x in out]


def test_survfit_without_a_gpu_is_a_loud_error():
    from jmbayes_b200 import api, build
    build.build_cuda()
    try:
        ndev = api.device_count()
    except api.JmbError:
        ndev = 0
    if ndev > 0:
        pytest.skip("a GPU is present")
    c = make_cohort("config2", n=12)
    d = sv.newdata_designs(c, L=2)
    with pytest.raises(api.JmbError):
        sv.SurvFit(d)
