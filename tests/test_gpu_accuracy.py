"""aucJM / prederrJM sums on the device (include/jmb200_accuracy.h) against the line-by-line restatement and through
size-independent properties at full size.  Sums of doubles in a different (fixed) order: 1e-11 relative.
"""
import numpy as np
import pytest

from oracle import oracle_accuracy as oa

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def acc(cuda_lib):
    from jmbayes_b200 import accuracy
    return accuracy


@pytest.mark.parametrize("n,seed,na", [(2, 1, 0.0), (300, 2, 0.0), (1500, 3, 0.0), (1500, 4, 0.15), (257, 5, 0.05)])
def test_auc_pairs_match_restated_loop(acc, n, seed, na):
    from test_accuracy_cpu import make_case
    case = make_case(n, seed, na)
    want = oa.aucJM_pairs(*case)
    got = acc.aucJM_pairs(*case)
    assert np.isclose(got["num"], want[1], rtol=1e-11) and np.isclose(got["den"], want[2], rtol=1e-11)
    assert (np.isnan(got["auc"]) and np.isnan(want[0])) or np.isclose(got["auc"], want[0], rtol=1e-11)


def test_auc_single_subject_is_na(acc):
    r = acc.aucJM_pairs([3.0], [1.0], 5.0, [0.4], [0.1], [0.2])
    assert np.isnan(r["auc"])


def test_auc_full_size_properties(acc):
    """200 000 subjects (2e10 pairs): the denominator against its O(n) prefix-sum form, and with nobody censored the AUC
    against the rank-based concordance."""
    rng = np.random.default_rng(11)
    n, Th = 200_000, 5.0
    Time = np.sort(rng.uniform(2.0, 9.0, n))
    event = rng.choice([0.0, 1.0, 2.0, 3.0], size=n, p=[0.35, 0.45, 0.1, 0.1])
    pi, pi2, pi3 = rng.uniform(size=n), rng.uniform(size=n), rng.uniform(size=n)
    r = acc.aucJM_pairs(Time, event, Th, pi, pi2, pi3)
    early, ev, ce = Time <= Th, np.isin(event, (1.0, 3.0)), np.isin(event, (0.0, 2.0))
    a = np.where(early & ev, 1.0, np.where(early & ce, pi2, 0.0))
    b = np.where(Time > Th, 1.0, np.where(early & ce, pi3, 0.0))
    den = float(np.sum(b * (np.cumsum(a) - a)))                    # sum_j b_j sum_{i<j} a_i
    assert np.isclose(r["den"], den, rtol=1e-11) and 0.0 < r["num"] < r["den"]
    ones = np.ones(n)
    r1 = acc.aucJM_pairs(Time, ones, Th, pi, pi2, pi3)
    cases, controls = pi[early], np.sort(pi[~early])
    conc = np.sum(controls.size - np.searchsorted(controls, cases, side="right")) / (cases.size * controls.size)
    assert r1["den"] == cases.size * controls.size and np.isclose(r1["auc"], conc, rtol=1e-12)


def test_prederr_sum_matches_restatement(acc):
    rng = np.random.default_rng(3)
    a, d, c, w = rng.uniform(size=3000), rng.uniform(size=2000), rng.uniform(size=1000), rng.uniform(size=1000)
    c[3] = np.nan
    for loss in ("square", "absolute"):
        assert np.isclose(acc.prederrJM_sum(a, d, c, w, 7000, loss), oa.prederrJM_sum(a, d, c, w, 7000, loss), rtol=1e-12)
    assert np.isclose(acc.prederrJM_sum(a, d, [], [], 5000), oa.prederrJM_sum(a, d, [], [], 5000), rtol=1e-12)


@pytest.mark.parametrize("n,seed,na", [(1, 1, 0.0), (700, 2, 0.0), (50_000, 3, 0.0), (300, 4, 0.05)])
def test_roc_sums_match_restated_outer_product(acc, n, seed, na):
    from test_accuracy_cpu import make_case
    Time, event, Th, pi, pi2, _ = make_case(n, seed, na)
    want = oa.rocJM(Time, event, Th, pi, pi2)
    got = acc.rocJM(Time, event, Th, pi, pi2)
    for k in ("nTP", "nFP", "nFN", "nTN", "TP", "FP"):
        assert np.allclose(got[k], want[k], rtol=1e-11, atol=1e-11, equal_nan=True), k
    # the kappa-type measures are 0/0-like where every (or no) probability is below the threshold: compare inside
    below = np.sum(np.asarray(pi)[:, None] < want["thrs"][None, :], axis=0)
    inside = (below > 0) & (below < np.sum(~np.isnan(pi)))
    for k in ("qSN", "qSP", "qOverall"):
        assert np.allclose(got[k][inside], want[k][inside], rtol=1e-9, atol=1e-9, equal_nan=True), k
    if not na:
        assert np.isclose(got["F1score"], want["F1score"], equal_nan=True) and np.isclose(got["Youden"], want["Youden"], equal_nan=True)
