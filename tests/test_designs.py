"""Design construction before the MCMC path (include/jmb200_designs.h): right_rows / right_rows_mstate and
splineDesign on the device.  CPU: the restatement against a brute-force scan and SciPy; GPU: bit-exact indices, 1e-14 bases."""
import numpy as np
import pytest

from oracle import oracle_designs as od


def _long_data(n, seed, K=15):
    rng = np.random.default_rng(seed)
    v = 1 + rng.poisson(4.0, size=n)
    ids = np.repeat(np.arange(100, 100 + n), v)                      # arbitrary labels, grouped by subject
    times = np.concatenate([np.sort(np.r_[0.0, rng.uniform(0, 8, k - 1)]) for k in v])
    T = rng.uniform(0.5, 10.0, size=n)
    Q = T[:, None] * rng.uniform(0, 1, size=(n, K))                  # quadrature-like points on (0, T)
    Q[0, 0] = -1.0                                                   # before the first visit: ind < 1 -> 1
    Q[1, :3] = times[ids == 101][:1]                                 # exactly at a visit time: that visit
    return ids, times, Q


def test_restated_right_rows_equals_brute_force():
    ids, times, Q = _long_data(300, 1)
    got = od.right_rows(times, ids, Q)
    first = np.flatnonzero(np.r_[True, ids[1:] != ids[:-1]])
    want = []
    for i in range(Q.shape[0]):
        lo, hi = first[i], (first[i + 1] if i + 1 < first.size else ids.size)
        for q in Q[i]:
            k = lo
            for r in range(lo, hi):
                if times[r] <= q:
                    k = r
            want.append(k)
    assert np.array_equal(got, np.array(want))
    idT = np.repeat(np.arange(300), 2)[:450].astype(np.int32)        # some subjects with two transition rows
    got2 = od.right_rows(times, ids, np.repeat(Q, 2, axis=0)[:450], idT=idT)
    assert np.array_equal(got2.reshape(450, -1), got.reshape(300, -1)[idT])


def test_restated_spline_design_properties():
    knots = np.r_[[0.0] * 4, np.linspace(0.5, 9.5, 13), [10.0] * 4]
    x = np.r_[np.linspace(0, 10, 501), -0.1, 10.2]
    D = od.spline_design(knots, x)
    assert D.shape == (503, 17) and np.allclose(D[:-2].sum(1), 1.0) and np.all(D[-2:] == 0.0)     # partition of unity, outer.ok
    assert np.all((D != 0).sum(1) <= 4) and D[500, -1] == 1.0                                       # band of ord, right-closed


def test_library_exports_the_design_entry_points():
    import ctypes as C
    from jmbayes_b200 import build
    L = C.CDLL(build.build_cuda())
    for name in ("jmb_right_rows", "jmb_right_rows_mstate", "jmb_spline_design", "jmb_data_rows"):
        assert hasattr(L, name), name


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(1, 3), (257, 4), (20000, 5)])
def test_right_rows_on_device_is_bit_exact(cuda_lib, n, seed):
    from jmbayes_b200 import designs
    ids, times, Q = _long_data(max(n, 2), seed)
    if n == 1:
        ids, times, Q = ids[ids == 100], times[ids == 100], Q[:1]
    assert np.array_equal(designs.right_rows(times, ids, Q), od.right_rows(times, ids, Q))
    nT = min(3 * Q.shape[0] // 2, 2 * Q.shape[0])
    idT = np.sort(np.r_[np.arange(Q.shape[0]), np.arange(nT - Q.shape[0])]).astype(np.int32)
    Qm = Q[idT] * 0.9
    assert np.array_equal(designs.right_rows(times, ids, Qm, idT=idT), od.right_rows(times, ids, Qm, idT=idT))
    if n > 1:
        bad = np.r_[ids, ids[:1]]                                    # the first subject's label again at the end: not grouped
        with pytest.raises(ValueError):
            designs.right_rows(np.r_[times, 9.0], bad, Q)


@pytest.mark.gpu
def test_spline_design_on_device_matches_scipy_and_the_host_path(cuda_lib):
    from jmbayes_b200 import designs
    rng = np.random.default_rng(2)
    knots = np.r_[[0.0] * 4, np.sort(rng.uniform(0.3, 9.7, 13)), [10.0] * 4]
    x = np.r_[rng.uniform(0, 10, 200000), 0.0, 10.0, knots[5], -0.5, 10.5]
    off, vals, D = designs.spline_design(knots, x, dense=True)
    want = od.spline_design(knots, x)
    assert np.max(np.abs(D - want)) <= 1e-14
    rows = np.arange(x.size)
    rebuilt = np.zeros_like(want)
    for r in range(4):
        rebuilt[rows, off + r] += vals[:, r]
    assert np.array_equal(rebuilt, D) and np.all(off >= 0) and np.all(off + 3 < 17)


def test_restated_data_rows_is_the_selected_frame_with_the_new_times():
    ids, times, Q = _long_data(50, 7, K=3)
    rng = np.random.default_rng(7)
    X = np.column_stack([times, rng.normal(size=times.size), (ids % 2).astype(float)])      # timeVar, a time-varying and a baseline covariate
    rows = od.right_rows(times, ids, Q)
    F = od.data_rows(X, rows, time_col=0, new_time=Q.ravel())
    assert F.shape == (150, 3) and np.array_equal(F[:, 0], Q.ravel()) and np.array_equal(F[:, 1:], X[rows][:, 1:])
    assert np.array_equal(od.data_rows(X, rows), X[rows])


@pytest.mark.gpu
def test_data_rows_on_device_is_bit_exact(cuda_lib):
    from jmbayes_b200 import designs
    ids, times, Q = _long_data(5000, 11)
    rng = np.random.default_rng(11)
    X = np.column_stack([times, rng.normal(size=times.size), rng.normal(size=times.size), (ids % 3).astype(float)])
    rows = designs.right_rows(times, ids, Q)
    got = designs.data_rows(X, rows, time_col=0, new_time=Q.ravel())
    assert np.array_equal(got, od.data_rows(X, rows, time_col=0, new_time=Q.ravel()))
    assert np.array_equal(designs.data_rows(X, rows), X[rows])
    with pytest.raises(Exception):
        designs.data_rows(X, np.r_[rows, X.shape[0]], time_col=0, new_time=np.r_[Q.ravel(), 1.0])     # row index out of range
