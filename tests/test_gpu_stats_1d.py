"""GPU parity tests (through the C-ABI / the MCSamples mirror): weighted moments, Gelman-Rubin, exact
weighted order statistics, 1D histograms and 1D densities against the oracle and the reference goldens."""
import numpy as np
import pytest

from cases import CASES, kw_tag
from helpers import load_case, make_oracle

pytestmark = pytest.mark.gpu

ALL = list(CASES)


def make_gpu(case, **kw):
    from getdist_b200 import MCSamples

    return MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"],
                     sampler=case.get("sampler", "uncorrelated"), settings=case["settings"] or None, **kw)


@pytest.fixture(scope="module")
def gpu_objs():
    cache = {}

    def get(name):
        if name not in cache:
            case, g = load_case(name)
            cache[name] = (case, g, make_gpu(case))
        return cache[name]

    return get


@pytest.mark.parametrize("name", ALL)
def test_moments(gpu_objs, name):
    case, g, mc = gpu_objs(name)
    assert np.max(np.abs(mc.getMeans() - g["means"]) / (np.abs(g["means"]) + np.sqrt(g["vars"]))) < 1e-12
    np.testing.assert_allclose(mc.getVars(), g["vars"], rtol=1e-12)
    scale = np.sqrt(np.outer(np.diag(g["cov"]), np.diag(g["cov"])))
    assert np.max(np.abs(mc.getCov() - g["cov"]) / scale) < 1e-12
    assert np.max(np.abs(mc.getCorrelationMatrix() - g["corr"])) < 1e-12
    np.testing.assert_allclose(mc.norm, float(g["norm"]), rtol=1e-13)
    assert mc.max_mult == float(g["max_mult"])
    np.testing.assert_allclose(mc.mean_mult, float(g["mean_mult"]), rtol=1e-13)


def test_gelman_rubin(gpu_objs):
    case, g, mc = gpu_objs("chains")
    np.testing.assert_allclose(mc.getGelmanRubin(), float(g["gelman_rubin"]), rtol=1e-9)
    np.testing.assert_allclose(mc.getGelmanRubin(3), float(g["gelman_rubin_3"]), rtol=1e-9)
    np.testing.assert_allclose(mc.getGelmanRubinEigenvalues(), g["gelman_rubin_eig"], rtol=1e-8, atol=1e-13)


@pytest.mark.parametrize("name", ALL)
def test_quantiles_bit_exact(gpu_objs, name):
    case, g, mc = gpu_objs(name)
    fr = g["quantile_fracs"]
    q = mc._ctx.weighted_quantiles(np.arange(mc.n), fr)
    assert np.array_equal(q, g["quantiles"]), np.max(np.abs(q - g["quantiles"]))
    # single-parameter / scalar forms of confidence()
    assert mc.confidence(0, 0.1) == g["quantiles"][0][2]
    assert np.array_equal(mc.twoTailLimits(1, 0.8), g["quantiles"][1][[2, 10]])


def test_quantiles_adversarial():
    """heavy duplicates, tight clusters, huge dynamic range, zero weights, a constant column"""
    from getdist_b200 import _abi
    from oracle.getdist_oracle import weighted_quantiles

    rng = np.random.default_rng(99)
    N = 300000
    cols = [
        rng.integers(0, 7, N).astype(np.float64),                       # 7 atoms
        np.where(rng.random(N) < 0.6, 1.25, rng.normal(size=N)),        # one atom holding 60% of the mass
        np.exp(rng.normal(0, 12, N)) * rng.choice([-1, 1], N),          # 10 orders of magnitude, both signs
        1e8 + 1e-6 * rng.normal(size=N),                                # cluster far from 0 (few distinct ulps)
        np.full(N, 3.5),                                                # constant
        np.round(rng.normal(size=N), 2),                                # many ties
        rng.normal(size=N),
    ]
    X = np.ascontiguousarray(np.stack(cols, axis=1))
    w = rng.integers(0, 4, N).astype(np.float64)  # integer weights incl. zeros: cumulative sums exact on both sides
    ctx = _abi.Context(0)
    ctx.set_samples(X, w)
    fr = np.array([0.0005, 0.001, 0.01, 0.1, 0.25, 0.5, 0.6, 0.75, 0.9, 0.99, 0.999, 0.9995])
    q = ctx.weighted_quantiles(np.arange(X.shape[1]), fr)
    for j in range(X.shape[1]):
        ref = weighted_quantiles(X[:, j], w, fr)
        assert np.array_equal(q[j], ref), (j, q[j], ref)
    ctx.close()


@pytest.mark.parametrize("name", ALL)
def test_hist1d_matches_bincount(gpu_objs, name):
    from oracle.getdist_oracle import bin_geometry, bin_indices

    case, g, mc = gpu_objs(name)
    o = make_oracle(case)
    mc._ensure_param_ranges(range(mc.n))
    specs = [mc._spec_1d(j, {}) for j in range(mc.n)]
    bins = mc._ctx.hist1d_batch(specs)
    for j in range(mc.n):
        par = o.init_param_ranges(j)
        binmin, binmax, fw = bin_geometry(par, 1024)
        # geometry agrees to rounding (it inherits the last-ulp difference of the device std dev) ...
        np.testing.assert_allclose([specs[j].binmin, specs[j].binmax], [binmin, binmax], rtol=1e-13)
        # ... and the histogram is compared on exactly the geometry the device was given
        binmin, binmax = specs[j].binmin, specs[j].binmax
        fw = (binmax - binmin) / (1024 - 1)
        ref = np.bincount(bin_indices(o.samples[:, j], binmin, fw), weights=o.weights, minlength=1024)
        assert np.all((ref == 0) == (bins[j] == 0))
        assert np.max(np.abs(bins[j] - ref)) <= 1e-11 * np.max(ref)


@pytest.mark.parametrize("name", ALL)
def test_density_1d(gpu_objs, name):
    case, g, mc = gpu_objs(name)
    for kw in case["kwargs_1d"]:
        tag = kw_tag(kw)
        for j in range(mc.n):
            d = mc.get1DDensityGridData(j, **kw)
            par = g["d1/%s/%d/par" % (tag, j)]
            p = mc.paramNames.names[j]
            got = np.array([p.range_min, p.range_max, p.sigma_range, p.param_min, p.param_max, p.err, p.mean,
                            float(p.has_limits_bot), float(p.has_limits_top)])
            np.testing.assert_allclose(got, par[:9], rtol=1e-12, atol=1e-12 * par[5])  # atol ~ sigma: the mean may be ~0
            if not np.isnan(par[9]):
                np.testing.assert_allclose(p.kde_h, par[9], rtol=2e-5)  # SURVEY s8c staged tolerance for h
            if len(par) > 11 and not np.isnan(par[11]):
                np.testing.assert_allclose(p.N_eff_kde, par[11], rtol=1e-9)  # incl. the mcmc autocorrelation estimate
            x = g["d1/%s/%d/x" % (tag, j)]
            assert d.x.size == int(x[2])
            np.testing.assert_allclose([d.x[0], d.x[-1]], x[:2], rtol=1e-13)
            err = np.max(np.abs(d.P - g["d1/%s/%d/P" % (tag, j)]))
            assert err < 1e-6, (name, tag, j, err)  # north-star bar
            assert err < 1e-9, (name, tag, j, err)  # what the implementation actually achieves


def test_mcmc_neff_and_corr_length(gpu_objs):
    """default sampler='mcmc': autocorrelation length by direct lag products == the reference's FFT route"""
    from oracle.getdist_oracle import correlation_length_rows, neff_mcmc

    case, g, mc = gpu_objs("mcmc")
    o = make_oracle(case)
    assert mc.sampler == "mcmc"
    for j in range(mc.n):
        x = o.samples[:, j]
        ref = correlation_length_rows(x, o.weights, o.means[j], o.vars[j])
        np.testing.assert_allclose(mc.getCorrelationLength(j, weight_units=False), ref, rtol=1e-9)
        np.testing.assert_allclose(mc.getEffectiveSamplesGaussianKDE(j, scale=0.37 * o.sddev[j]),
                                   neff_mcmc(x, o.weights, 0.37 * o.sddev[j]), rtol=1e-9)


def test_density_1d_cache_and_names(gpu_objs):
    case, g, mc = gpu_objs("mix3")
    d = mc.get1DDensity("a")
    assert mc.get1DDensity("a") is d          # cached (mcsamples.py:1510-1513)
    assert mc.get1DDensity("a", fine_bins=512) is not d  # kwargs bypass the cache
    assert mc.get1DDensity("nope") is None   # unknown name -> None (mcsamples.py:1538-1539)


def test_larger_n_properties():
    """size-independent properties at a larger N: histogram mass conservation, mirror symmetry of the density
    of mirrored samples, idempotent re-upload"""
    from getdist_b200 import MCSamples

    rng = np.random.default_rng(3)
    N = 2_000_000
    x = rng.normal(size=(N, 2)) * np.array([1.0, 3.0]) + np.array([0.0, 7.0])
    w = rng.exponential(1.0, N)
    mc = MCSamples(samples=x, weights=w, names=["u", "v"], sampler="uncorrelated", settings={"fine_bins": 2048})
    mc._ensure_param_ranges([0, 1])
    bins = mc._ctx.hist1d_batch([mc._spec_1d(j, {}) for j in range(2)])
    np.testing.assert_allclose(bins.sum(axis=1), w.sum(), rtol=1e-11)
    d = mc.get1DDensity("u")
    mc2 = MCSamples(samples=-x, weights=w, names=["u", "v"], sampler="uncorrelated", settings={"fine_bins": 2048})
    d2 = mc2.get1DDensity("u")
    assert np.allclose(d.P, d2.P[::-1], atol=1e-10)
    assert np.allclose(d.x, -d2.x[::-1])
    np.testing.assert_allclose(mc.getMeans(), -mc2.getMeans(), rtol=1e-13)


def test_pickle_and_copy_roundtrip(gpu_objs):
    """the object stays picklable / deep-copyable (device handle excluded, samples re-uploaded lazily)"""
    import pickle

    case, g, mc = gpu_objs("mix3")
    ref = mc.get1DDensity("b").P.copy()
    mc2 = pickle.loads(pickle.dumps(mc))
    assert np.array_equal(mc2.get1DDensity("b").P, ref)
    np.testing.assert_allclose(mc2.getCov(), mc.getCov(), rtol=0, atol=0)
    mc3 = mc.copy(settings={"fine_bins": 512})
    assert mc3.get1DDensity("b").P.size == 512 and mc.get1DDensity("b").P.size == 1024


@pytest.mark.parametrize("name", ["bounded", "likes"])
def test_marge_limits(name):
    """MCSamples.setMargeLimits (batched 1D densities + one device quantile call + the host limit logic)
    against the limits of the unmodified reference (tests/golden/limits.npz)"""
    import os

    from getdist_b200 import MCSamples
    from helpers import GOLDEN

    g = np.load(os.path.join(GOLDEN, "limits.npz"))
    case, _ = load_case(name)
    mc = MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"],
                   sampler="uncorrelated", settings=case["settings"] or None)
    lims = mc.setMargeLimits()
    tags = {0: "two", 1: ">", 2: "<", 3: "none"}
    for j, nm in enumerate(case["names"]):
        ref, rt = g["%s/%d/limits" % (name, j)], g["%s/%d/tags" % (name, j)]
        assert [l.limitTag() for l in lims[nm]] == [tags[int(t)] for t in rt], (name, nm)
        got = np.array([[l.lower, l.upper] for l in lims[nm]], dtype=np.float64)
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7 * mc.sddev[j], err_msg=str((name, nm)))
