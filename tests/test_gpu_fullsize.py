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
        assert fit.layout_info()["kernel"] == "static-tma:three_outcomes_ic"
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
    # the direct (non-staged) kernel must walk the same chain
    monkeypatch.setenv("JMB_KERNEL", "direct")
    with cuda_lib.Fit(D, c["Covs"]["b"], True) as fit:
        fit.set_draw(D)
        chain2 = fit.lap_rwm_C(ini, c["priors"], c["scales"], c["Covs"], ctl)
    assert np.allclose(chain2["extras"]["trace_surv"], tr, rtol=1e-12)
    assert np.array_equal(chain2["mcmc"]["b"], chain["mcmc"]["b"])
    assert abs(tr[0] - (got["log_ptb"].sum())) < 0.05 * abs(tr[0])      # one b-step away from the initial state
