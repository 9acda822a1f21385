"""aucJM / prederrJM sums: the line-by-line restatement (oracle/oracle_accuracy.py) against an independent evaluation of the
same sums (the pair weight factorises into a_i b_j, include/jmb200_accuracy.h), and the C ABI exports."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle_accuracy as oa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_case(n, seed, na_frac=0.0, Thoriz=5.0):
    rng = np.random.default_rng(seed)
    Time = np.sort(rng.uniform(2.0, 9.0, n))                       # newdata2 is sorted by Time, all > Tstart
    event = rng.choice([0.0, 1.0, 2.0, 3.0], size=n, p=[0.35, 0.45, 0.1, 0.1])
    pi = rng.uniform(0.05, 0.95, n)
    pi[::7] = 0.5                                                  # ties: pi_i < pi_j is strict
    pi2, pi3 = rng.uniform(0.05, 0.95, n), rng.uniform(0.05, 0.95, n)
    if na_frac:
        for v in (pi, pi2, pi3):
            v[rng.uniform(size=n) < na_frac] = np.nan
    return Time, event, Thoriz, pi, pi2, pi3


def factorised(Time, event, Thoriz, pi, pi2, pi3):
    """Same sums from per-subject factors a_i, b_j with explicit double loops over the classes (small n only)."""
    n = Time.size
    early = Time <= Thoriz
    ev, ce = np.isin(event, (1.0, 3.0)), np.isin(event, (0.0, 2.0))
    a = np.where(early & ev, 1.0, np.where(early & ce, pi2, 0.0))
    b = np.where(Time > Thoriz, 1.0, np.where(early & ce, pi3, 0.0))
    num = den = 0.0
    for i in range(n):
        if a[i] == 0.0:
            continue
        for j in range(i + 1, n):
            if b[j] == 0.0:
                continue
            w = a[i] * b[j]
            if np.isnan(w):
                continue
            den += w
            if not (np.isnan(pi[i]) or np.isnan(pi[j])) and pi[i] < pi[j]:
                num += w
    return (num / den if den else float("nan")), num, den


@pytest.mark.parametrize("n,seed,na", [(2, 1, 0.0), (40, 2, 0.0), (150, 3, 0.0), (150, 4, 0.15)])
def test_restated_pair_loop_equals_factorised_sums(n, seed, na):
    case = make_case(n, seed, na)
    auc, num, den = oa.aucJM_pairs(*case)
    auc2, num2, den2 = factorised(*case)
    assert np.isclose(num, num2, rtol=1e-12) and np.isclose(den, den2, rtol=1e-12)
    assert (np.isnan(auc) and np.isnan(auc2)) or np.isclose(auc, auc2, rtol=1e-12)


def test_auc_is_the_concordance_when_nobody_is_censored():
    rng = np.random.default_rng(9)
    n = 120
    Time = np.sort(rng.uniform(1, 10, n)); event = np.ones(n); pi = rng.uniform(size=n)
    auc, num, den = oa.aucJM_pairs(Time, event, 5.0, pi, np.zeros(n), np.zeros(n))
    cases, controls = pi[Time <= 5.0], pi[Time > 5.0]
    assert den == cases.size * controls.size
    assert np.isclose(auc, np.mean(cases[:, None] < controls[None, :]))
    assert np.isnan(oa.aucJM_pairs(Time[:1], event[:1], 5.0, pi[:1], pi[:1], pi[:1])[0])


def test_prederr_sum_restates_the_three_groups():
    rng = np.random.default_rng(3)
    a, d, c, w = rng.uniform(size=30), rng.uniform(size=20), rng.uniform(size=10), rng.uniform(size=10)
    c[3] = np.nan
    for loss, L in (("square", np.square), ("absolute", np.abs)):
        want = (np.sum(L(1 - a)) + np.sum(L(d)) + np.nansum(w * L(1 - c) + (1 - w) * L(c))) / 77
        assert np.isclose(oa.prederrJM_sum(a, d, c, w, 77, loss), want, rtol=1e-13)


def test_cabi_library_exports_the_accuracy_symbols():
    from jmbayes_b200 import build
    lib = build.build_cuda()
    hdr = open(os.path.join(ROOT, "include", "jmb200_accuracy.h")).read()
    names = sorted(set(re.findall(r"\b(jmb_[A-Za-z_0-9]+)\s*\(", hdr)))
    assert names == ["jmb_aucJM_pairs", "jmb_prederrJM_sum", "jmb_rocJM_sums"]
    L = C.CDLL(lib)
    for nm in names:
        assert hasattr(L, nm), nm


def test_roc_restatement_on_a_case_worked_by_hand():
    # three subjects: one dead before the horizon (pi .2), one censored before it (pi .6, weight .5), one alive after it (pi .9)
    r = oa.rocJM([1.0, 2.0, 9.0], [1.0, 0.0, 0.0], 5.0, [0.2, 0.6, 0.9], [0.0, 0.5, 0.0])
    k = 50                                                         # threshold 0.5: only the first subject is below it
    assert np.isclose(r["nTP"][k], 1.0) and np.isclose(r["nFP"][k], 0.0) and np.isclose(r["nFN"][k], 0.5)
    assert np.isclose(r["TP"][k], 1.0 / 1.5) and np.isclose(r["nTN"][k], 1.5)
    assert np.isclose(r["nTP"][100], 1.5) and np.isclose(r["nFP"][100], 1.5)      # threshold 1: everybody below
