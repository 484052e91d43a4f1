"""The reference's own known-answer tests, run against the oracle restatement
(reference test file:line cited per test)."""
import numpy as np
import pytest

from oracle import diagnostics as od
from oracle import samplers as osm
from oracle.models import IsoGauss


def test_autocorr_fixed():  # test_autocorr.py:10-14
    ac = od.autocorr(np.asarray([1, 0, 0, 0]))
    np.testing.assert_allclose([1.000, -0.083, -0.167, -0.250], ac, atol=0.001, rtol=0.001)


def test_autocorr_exceptions():  # test_autocorr.py:42-47
    for bad in ([], [1.1]):
        with pytest.raises(ValueError):
            od.autocorr(bad)
    od.autocorr([1.1, 1.2])


@pytest.mark.parametrize("chain,pos", [  # test_iat.py:72-80
    ([], 0), ([1], 0), ([1, -0.5], 2), ([1, -0.5, 0.25], 2), ([1, -0.5, 0.25, -0.3], 2),
    ([1, -0.5, 0.25, -0.1], 4), ([1, -0.5, 0.25, -0.3, 0.05], 2), ([1, -0.5, 0.25, -0.1, 0.05], 4)])
def test_end_pos_pairs(chain, pos):
    assert od.end_pos_pairs(chain) == pos


def _iat_ar1(phi):  # test_iat.py:18-26
    return (1 + phi) / (1 - phi)


@pytest.mark.parametrize("phi", [-0.5, -0.3, -0.1, 0.1, 0.3])
def test_iat_ess_ar1(phi):  # test_iat.py:29-54, test_ess.py:14-39
    rng = np.random.default_rng(int(100 + 10 * phi))
    v = od.sample_ar1(phi, 20_000, rng)
    for f in (od.iat, od.iat_imse, od.iat_ipse):
        np.testing.assert_allclose(_iat_ar1(phi), f(v), rtol=0.1)
    for f in (od.ess, od.ess_imse, od.ess_ipse):
        np.testing.assert_allclose(20_000 / _iat_ar1(phi), f(v), rtol=0.1)


def test_short_chain_exceptions():  # test_iat.py:57-65, test_ess.py:42-50
    for n in range(4):
        v = np.arange(n, dtype=float)
        for f in (od.iat, od.iat_imse, od.iat_ipse, od.ess, od.ess_imse, od.ess_ipse):
            with pytest.raises(ValueError):
                f(v)


def _rhat_bda3(chains):  # test_rhat.py:19-31
    chains = [np.asarray(c) for c in chains]
    m = np.array([c.mean() for c in chains])
    M, N = len(chains), len(chains[0])
    B = N / (M - 1) * np.sum((m - m.mean()) ** 2)
    W = np.mean([np.sum((c - c.mean()) ** 2) / (N - 1) for c in chains])
    return np.sqrt(((N - 1) / N * W + B / N) / W)


def test_rhat_vs_bda3():  # test_rhat.py:34-52
    c = [[1.01, 1.05, 0.98, 0.90, 1.23], [0.99, 1.00, 1.01, 1.15, 0.83],
         [0.84, 0.90, 0.94, 1.10, 0.92], [0.32, 1.81, 0.90, 0.10, 2.85]]
    for k in (2, 3, 4):
        np.testing.assert_allclose(_rhat_bda3(c[:k]), od.rhat(c[:k]), rtol=1e-12)


def test_rhat_exceptions():  # test_rhat.py:55-68
    for bad in ([], [[1.01, 1.2, 1.3, 1.4]], [[1, 2, 3], [4], [5, 6, 7, 8, 9]]):
        with pytest.raises(ValueError):
            od.rhat(bad)


def test_split_chains():  # test_rhat.py:71-79
    got = od.split_chains([[1, 2, 3], [4, 5, 6, 7]])
    assert [list(g) for g in got] == [[1, 2], [3], [4, 5], [6, 7]]
    np.testing.assert_allclose(od.rhat([[1, 2], [3, 4]]), od.split_rhat([[1, 2, 3, 4]]))


def test_hmc_one_step_is_mala():  # test_equivalencies.py:12-32
    model = IsoGauss(1)
    rng = np.random.default_rng(123)
    z = rng.standard_normal((50, 1)); u = rng.random(50)
    eps = 0.02
    d1, _, _ = osm.hmc_diag(model, np.array([0.2]), z, u, eps, 1)
    d2, _, _ = osm.mala(model, np.array([0.2]), z, u, 0.5 * eps ** 2)
    np.testing.assert_array_almost_equal(d1, d2)
    assert len(np.unique(d1)) > 20


def test_accept_rule():  # test_metropolis.py:32-52 (uniform draw 0.5)
    model = IsoGauss(1)
    for p_cur, expect in ((0.81, False), (0.79, True)):
        th0 = np.array([np.sqrt(-2 * np.log(p_cur))])
        step = np.sqrt(-2 * np.log(0.4)) - th0[0]      # proposal density 0.4
        _, _, a = osm.metropolis_rw(model, th0, np.array([[step]]), np.array([0.5]), 1.0)
        assert bool(a[0]) is expect
    # log(0) = -inf always accepts (metropolis.py:36-38)
    _, _, a = osm.metropolis_rw(model, np.array([0.1]), np.array([[5.0]]), np.array([0.0]), 1.0)
    assert a[0]


# ---- rank-normalised R-hat family -------------------------------------
def _rank_norm(r, S):  # test_rhat.py:133-134
    from scipy.stats import norm
    return norm.ppf((r - 0.325) / (S - 0.25))


def test_split_chains():  # test_rhat.py:71-79
    eq = lambda want, got: [np.testing.assert_array_equal(w, g) for w, g in zip(want, got)] and None
    assert od.split_chains([]) == []
    eq([[1], []], od.split_chains([[1]]))
    eq([[1], [2]], od.split_chains([[1, 2]]))
    eq([[1, 2], [3]], od.split_chains([[1, 2, 3]]))
    eq([[1, 2], [3], [4, 5], [6, 7]], od.split_chains([[1, 2, 3], [4, 5, 6, 7]]))


def test_split_rhat():  # test_rhat.py:82-110
    np.testing.assert_allclose(od.rhat([[1, 2], [3, 4]]), od.split_rhat([[1, 2, 3, 4]]))
    np.testing.assert_allclose(od.rhat([[1, -2, 3], [4, 5, 6], [7, 8], [9, 12]]),
                               od.split_rhat([[1, -2, 3, 4, 5, 6], [7, 8, 9, 12]]))
    for bad in ([], [[1, 2, 3]], [[1, 2, 3, 4], [1, 2, 3]]):
        with pytest.raises(ValueError):
            od.split_rhat(bad)


def test_rank_chains():  # test_rhat.py:113-126
    assert od.rank_chains([]) == []
    for want, chains in [([[1]], [[2.3]]), ([[2, 3, 1]], [[3.9, 5.2, 2.1]]), ([[2], [1]], [[4.2], [1.9]]),
                         ([[2, 3], [5, 4], [1, 6]], [[4.2, 5.7], [7.2, 6.1], [-12.9, 107]])]:
        for w, g in zip(want, od.rank_chains(chains)):
            np.testing.assert_array_equal(w, g)


def test_rank_normalize_and_rhat():  # test_rhat.py:137-196
    np.testing.assert_array_equal([[_rank_norm(1, 1)]], od.rank_normalize_chains([[32.7]]))
    got = od.rank_normalize_chains([[3.9, 3.1], [2.2, 5.9]])
    np.testing.assert_array_equal([[_rank_norm(3, 4), _rank_norm(2, 4)], [_rank_norm(1, 4), _rank_norm(4, 4)]], got)
    rn = [_rank_norm(i, 8) for i in range(1, 9)]
    np.testing.assert_allclose(od.split_rhat([[rn[1], rn[2], rn[6], rn[7]], [rn[0], rn[3], rn[5], rn[4]]]),
                               od.rank_normalized_rhat([[2, 3, 7, 8], [1, 4, 6, 5]]))
    for bad in ([], [[1.01, 1.2, 1.3]], [[1, 2, 3], [4]]):
        with pytest.raises(ValueError):
            od.rank_normalized_rhat(bad)
