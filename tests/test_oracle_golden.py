"""Pins the CPU oracle (oracle/getdist_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from cases import CASES, grid_stride, kw_tag
from helpers import load_case, make_oracle

ALL = list(CASES)


@pytest.mark.parametrize("name", ALL)
def test_moments(name):
    case, g = load_case(name)
    o = make_oracle(case)
    assert np.max(np.abs(o.get_means() - g["means"]) / (np.abs(g["means"]) + np.sqrt(g["vars"]))) < 1e-13
    np.testing.assert_allclose(o.get_vars(), g["vars"], rtol=1e-12)
    scale = np.sqrt(np.outer(np.diag(g["cov"]), np.diag(g["cov"])))
    assert np.max(np.abs(o.get_cov() - g["cov"]) / scale) < 1e-13
    assert np.max(np.abs(o.get_correlation_matrix() - g["corr"])) < 1e-13
    assert float(o.norm) == float(g["norm"])
    assert float(o.max_mult) == float(g["max_mult"])


def test_gelman_rubin():
    case, g = load_case("chains")
    o = make_oracle(case)
    assert np.array_equal(o.chain_offsets, g["chain_offsets"])
    np.testing.assert_allclose(o.get_gelman_rubin(), float(g["gelman_rubin"]), rtol=1e-10)
    np.testing.assert_allclose(o.get_gelman_rubin(3), float(g["gelman_rubin_3"]), rtol=1e-10)


@pytest.mark.parametrize("name", ALL)
def test_quantiles_exact(name):
    from oracle.getdist_oracle import weighted_quantiles

    case, g = load_case(name)
    o = make_oracle(case)
    for j in range(o.n):
        q = weighted_quantiles(o.samples[:, j], o.weights, g["quantile_fracs"])
        assert np.array_equal(q, g["quantiles"][j])


@pytest.mark.parametrize("name", ALL)
def test_density_1d(name):
    case, g = load_case(name)
    o = make_oracle(case)
    for kw in case["kwargs_1d"]:
        tag = kw_tag(kw)
        for j in range(o.n):
            d = o.density_1d(j, **kw)
            par = g["d1/%s/%d/par" % (tag, j)]
            p = o.pars[j]
            got = np.array([p.range_min, p.range_max, p.sigma_range, p.param_min, p.param_max, p.err, p.mean,
                            float(p.has_limits_bot), float(p.has_limits_top)])
            np.testing.assert_allclose(got, par[:9], rtol=1e-13, atol=1e-13 * par[5])  # atol ~ sigma: the mean may be ~0
            if not np.isnan(par[9]):
                np.testing.assert_allclose(p.kde_h, par[9], rtol=1e-9)
            if len(par) > 11 and not np.isnan(par[11]):
                np.testing.assert_allclose(p.N_eff_kde, par[11], rtol=1e-9)
            x = g["d1/%s/%d/x" % (tag, j)]
            assert d.x.size == int(x[2])
            np.testing.assert_allclose([d.x[0], d.x[-1]], x[:2], rtol=1e-14)
            assert np.max(np.abs(d.P - g["d1/%s/%d/P" % (tag, j)])) < 1e-10, (name, tag, j)


@pytest.mark.parametrize("name", ALL)
def test_density_2d(name):
    case, g = load_case(name)
    o = make_oracle(case)
    for kw in case["kwargs_2d"]:
        tag = kw_tag(kw)
        for (jx, jy) in case["pairs"]:
            d = o.density_2d(jx, jy, **kw)
            xy = g["d2/%s/%d_%d/xy" % (tag, jx, jy)]
            assert d.x.size == int(xy[2]) and d.y.size == int(xy[5])
            np.testing.assert_allclose([d.x[0], d.x[-1], d.y[0], d.y[-1]], xy[[0, 1, 3, 4]], rtol=1e-14)
            st = grid_stride(d.P.shape[0])
            ref = g["d2/%s/%d_%d/P" % (tag, jx, jy)]
            assert np.max(np.abs(d.P[::st, ::st] - ref)) < 1e-9, (name, tag, jx, jy)


def test_meanlikes_and_mask_function():
    """mean likelihoods (mcsamples.py:1556-1561, 1672-1684, 1829-1831, 1886-1901, 2004-2006) and
    mask_function (mcsamples.py:1909-1919, 1973-1979) of the oracle against the reference goldens"""
    case, g = load_case("likes")
    o = make_oracle(case)
    np.testing.assert_allclose(o.mean_loglike, float(g["mean_loglike"]), rtol=1e-13)
    for kw in case["likes_kwargs_1d"]:
        tag = kw_tag(kw)
        for j in range(o.n):
            d = o.density_1d(j, meanlikes=True, **kw)
            assert np.max(np.abs(d.P - g["l1/%s/%d/P" % (tag, j)])) < 1e-10
            assert np.max(np.abs(d.likes - g["l1/%s/%d/likes" % (tag, j)])) < 1e-9, (tag, j)
    for kw in case["likes_kwargs_2d"]:
        tag = kw_tag(kw)
        for (jx, jy) in case["pairs"]:
            d = o.density_2d(jx, jy, meanlikes=True, **kw)
            assert np.max(np.abs(d.P - g["l2/%s/%d_%d/P" % (tag, jx, jy)])) < 1e-9
            assert np.max(np.abs(d.likes - g["l2/%s/%d_%d/likes" % (tag, jx, jy)])) < 1e-8, (tag, jx, jy)
    for kw in case["mask_kwargs_2d"]:
        tag = kw_tag(kw)
        for (jx, jy) in case["mask_pairs"]:
            d = o.density_2d(jx, jy, mask_function=case["mask_function"], **kw)
            assert np.array_equal(d.mask, g["m2/%s/%d_%d/mask" % (tag, jx, jy)])
            assert np.max(np.abs(d.P - g["m2/%s/%d_%d/P" % (tag, jx, jy)])) < 1e-9, (tag, jx, jy)
