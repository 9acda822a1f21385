"""BASELINE.json full size (config 4: 1M subjects, 3 outcomes, interval censoring + left
truncation) on one GPU, checked through size-independent properties: agreement with the oracle on
random sub-cohorts, checksum of checksums (per-subject pieces vs the chain's per-iteration sums),
and identical chains from the TMA-staged and the direct kernels."""
import os

import numpy as np
import pytest

from jmbayes_b200.cohorts import make_cohort, shared_model

pytestmark = pytest.mark.gpu
N_FULL = int(os.environ.get("JMB_FULL_N", "1000000"))


def _slice_subjects(c, idx):
    """Sub-cohort made of the (sorted) subjects ``idx``."""
    D = c["Data"]
    n = int(np.asarray(D["event"]).shape[0])
    K = int(np.asarray(D["Pw"]).shape[0]) // n
    idx = np.asarray(idx)
    q = (idx[:, None] * K + np.arange(K)[None, :]).reshape(-1)
    out = dict(D)
    ys, Xb, Zs, idLs, idL2s = [], [], [], [], []
    for k in range(len(D["y"])):
        idL = np.asarray(D["idL"][k])
        last = np.asarray(D["idL2"][k]).astype(np.int64)
        first = np.concatenate([[0], last[:-1] + 1])
        rows = np.concatenate([np.arange(first[i], last[i] + 1) for i in idx])
        cnt = (last[idx] - first[idx] + 1)
        ys.append(np.asarray(D["y"][k])[rows]); Xb.append(np.asarray(D["Xbetas"][k])[rows])
        Zs.append(np.asfortranarray(np.asarray(D["Z"][k])[rows]))
        idLs.append(np.repeat(np.arange(idx.size), cnt).astype(np.int32))
        idL2s.append((np.cumsum(cnt) - 1).astype(np.int32))
    out.update(y=ys, Xbetas=Xb, Z=Zs, idL=idLs, idL2=idL2s)
    for key in ("event", "W1", "W2"):
        out[key] = np.asfortranarray(np.asarray(D[key])[idx])
    for key in ("W1s", "W2s", "Pw", "W1s_int", "W2s_int", "Pw_int"):
        out[key] = np.asfortranarray(np.asarray(D[key])[q])
    for key in ("XXbetas", "ZZ", "U"):
        out[key] = [np.asfortranarray(np.asarray(x)[idx]) for x in D[key]]
    for key in ("XXsbetas", "ZZs", "Us", "XXsbetas_int", "ZZs_int", "Us_int"):
        out[key] = [np.asfortranarray(np.asarray(x)[q]) for x in D[key]]
    out["idGK_fast"] = (np.arange(1, idx.size + 1) * K - 1).astype(np.int32)
    return out, np.asfortranarray(np.asarray(c["Covs"]["b"])[:, :, idx])


def test_full_size_config4_properties(cuda_lib, oracle, monkeypatch):
    shared = shared_model("config4", N_FULL)
    c = make_cohort("config4", n=N_FULL, seed=424242, shared=shared)
    D, ini = c["Data"], c["inits"]
    ctl = dict(c["control"]); ctl.update(n_burnin=2, n_iter=2, n_block=2, n_thin=2)
    with cuda_lib.Fit(D, c["Covs"]["b"], True) as fit:
        assert fit.layout_info()["kernel"].startswith("static-rows") and fit.layout_info()["kernel"].endswith(":three_outcomes_ic")
        fit.set_draw(D)
        got = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        rp = fit.row_ptr(2)
        assert rp[-1] == len(D["y"][2]) and np.array_equal(rp[1:] - 1, np.asarray(D["idL2"][2], dtype=np.int64))
        chain = fit.lap_rwm_C(ini, c["priors"], c["scales"], c["Covs"], ctl)
    for key in ("log_pyb", "log_pb", "log_h", "H", "HL", "log_ptb"):
        assert np.isfinite(got[key]).all(), key
    # checksum of checksums: the first iteration's "current" sum is the sum of the per-subject pieces
    # after that iteration's b-step; it must lie between plausible bounds and be reproducible
    tr = chain["extras"]["trace_surv"]
    assert np.isfinite(tr).all() and tr.shape == (8,)
    # random sub-cohorts against the oracle (indexing at scale: first, middle, last subjects)
    rng = np.random.default_rng(3)
    for lo in (0, N_FULL // 2, N_FULL - 400):
        idx = np.sort(rng.choice(np.arange(lo, lo + 400), size=120, replace=False))
        subD, subR = _slice_subjects(c, idx)
        oracle.set_rowsum_mode(1)
        ref = oracle.eval_logpost_re(subD, subR, True, np.asarray(ini["b"])[idx], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        for key in ("log_pyb", "log_pb", "log_h", "H", "HL", "log_ptb"):
            err = np.max(np.abs(got[key][idx] - ref[key]) / np.maximum(np.abs(ref[key]), 1e-12))
            assert err <= 1e-10, (key, lo, float(err))
    # the one-lane-per-row kernels (TMA-staged, direct) must walk the same chain: same accept decisions, b to 1e-9
    for variant in ("tma", "direct"):
        monkeypatch.setenv("JMB_KERNEL", variant)
        with cuda_lib.Fit(D, c["Covs"]["b"], True) as fit:
            fit.set_draw(D)
            chain2 = fit.lap_rwm_C(ini, c["priors"], c["scales"], c["Covs"], ctl)
        assert np.allclose(chain2["extras"]["trace_surv"], tr, rtol=1e-11), variant
        assert np.allclose(chain2["mcmc"]["b"], chain["mcmc"]["b"], rtol=1e-9, atol=1e-12), variant
    assert abs(tr[0] - (got["log_ptb"].sum())) < 0.05 * abs(tr[0])      # one b-step away from the initial state


# ---- BASELINE config 5 at the size of one GPU's shard: 125 000 new subjects, M = 200 draws, L = 10 horizons --------------
N_SVFT = int(os.environ.get("JMB_FULL_SVFT_N", "125000"))


def test_survfit_full_shard_properties(cuda_lib):
    """Size-independent properties of survfitJM on the full per-GPU shard of config 5, and agreement with the restated R
    loop on a random sub-cohort (draws are keyed by the global subject index, so a sub-cohort reproduces its subjects)."""
    from jmbayes_b200 import survfit as sv
    from oracle import oracle_svft
    c = make_cohort("config3", n=N_SVFT, seed=777)
    means, draws = sv.posterior_draws(c, M=200)
    d = sv.newdata_designs(c, L=10, dense_basis=False)           # bases evaluated by the library from (knots, times)
    with sv.SurvFit(d) as h:
        r = h.survfitJM(means, draws, full=False)
        info = h.info()
    S = r["summaries"]                                            # L x 4 x n: Mean, Median, Lower, Upper
    assert S.shape == (10, 4, N_SVFT) and np.all(np.isfinite(S))
    assert np.all(S > 0) and np.all(S <= 1.0)
    assert np.all(S[:, 2] <= S[:, 1] + 1e-15) and np.all(S[:, 1] <= S[:, 3] + 1e-15)     # Lower <= Median <= Upper
    assert np.all(np.diff(S[:, 0], axis=0) <= 1e-15)              # the mean survival curve decreases over the horizons
    assert np.mean(r["convergence"] == 0) > 0.999 and np.all(np.isfinite(r["modes"]))
    assert 0.3 < r["accept_rate"].mean() < 0.8 and r["accept_rate"].min() > 0.02
    assert info["launches"] == 4                                  # mode search, sampling, horizons, summaries
    # random sub-cohort against the oracle: same subjects, same global indices
    rng = np.random.default_rng(5)
    for start in rng.integers(0, N_SVFT - 8, size=3):
        sel = np.arange(start, start + 8)
        ds = sv.newdata_designs(c, L=10, subjects=sel)
        o = oracle_svft.survfitJM(ds, means, draws, subject_offset=int(start))
        assert np.max(np.abs(o["modes"] - r["modes"][sel])) < 1e-6
        diff = np.max(np.abs(o["summaries"] - S[:, :, sel]), axis=(0, 1))
        assert np.sum(diff > 1e-5) <= 1                           # an MH decision can flip only within ~1e-7 of the ratio
