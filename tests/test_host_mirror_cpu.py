"""Host mirror (getdist_b200/mcsamples.py) on the CPU: the ctypes Context is replaced by a TEST DOUBLE that answers
the same calls with numpy (moments, exact weighted quantiles) and with the device grid-stage code compiled for the
host (tests/hostsim: kde1d_core on np.bincount histograms).  Checks the scalar host logic -- _initParam ranges,
spec building, caching, marginalised limits -- without a GPU.  The product path itself never runs like this: the
real Context raises without libgdk.so / a CUDA device."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import GOLDEN, load_case, make_oracle
from test_hostsim import Res1D, Spec1D, dptr, hs  # noqa: F401  (hs: session fixture building tests/hostsim)


class FakeContext:
    """numpy / hostsim stand-in for getdist_b200._abi.Context (1D path only)"""

    hs = None

    def __init__(self, device=0):
        self.device = device

    def close(self):
        pass

    def set_samples(self, X, w=None, chain_offsets=None):
        self.X = np.asarray(X, dtype=np.float64)
        self.N, self.P = self.X.shape
        self.w = np.ones(self.N) if w is None else np.asarray(w, dtype=np.float64)
        self.offs = np.array([0, self.N]) if chain_offsets is None else np.asarray(chain_offsets)
        self.nchains = len(self.offs) - 1

    def moments(self):
        from oracle.getdist_oracle import weighted_cov, weighted_means, weighted_vars

        m = weighted_means(self.X, self.w)
        cm, cc, cn = [], [], []
        for a, b in zip(self.offs[:-1], self.offs[1:]):
            mm = weighted_means(self.X[a:b], self.w[a:b])
            cm.append(mm)
            cc.append(weighted_cov(self.X[a:b], self.w[a:b], mm))
            cn.append(self.w[a:b].sum())
        return dict(means=m, vars=weighted_vars(self.X, self.w, m), cov=weighted_cov(self.X, self.w, m),
                    scalars=np.array([self.w.sum(), (self.w**2).sum(), self.w.max(), 0.0, self.N, self.w.min(), 0, 0]),
                    xmin=self.X.min(axis=0), xmax=self.X.max(axis=0), chain_means=np.array(cm),
                    chain_covs=np.array(cc), chain_norms=np.array(cn))

    def weighted_quantiles(self, params, fracs):
        from oracle.getdist_oracle import weighted_quantiles

        return np.array([weighted_quantiles(self.X[:, j], self.w, np.asarray(fracs)) for j in params])

    def weighted_quantiles_range(self, params, fracs, start, end):
        from oracle.getdist_oracle import weighted_quantiles

        return np.array([weighted_quantiles(self.X[start:end, j], self.w[start:end], np.asarray(fracs)) for j in params])

    def weight_fraction_rows(self, fracs):
        return np.searchsorted(np.cumsum(self.w), np.asarray(fracs) * self.w.sum())

    def set_loglikes(self, ll):
        self.ll = np.asarray(ll, dtype=np.float64)
        return float(self.w.dot(self.ll) / self.w.sum())

    def histnd(self, params, nbins, binmin, binmax, which=0):
        from oracle.getdist_oracle import bin_indices

        flat = np.zeros(self.N, dtype=np.int64)
        stride = 1
        for j, n, lo, hi in zip(params, nbins, binmin, binmax):
            flat += bin_indices(self.X[:, j], lo, (hi - lo) / (n - 1)) * stride
            stride *= n
        shape = tuple(int(n) for n in nbins[::-1])
        if which == 2:
            out = np.zeros(stride)
            np.maximum.at(out, flat, np.exp(self.ll.min() - self.ll))
            return out.reshape(shape)
        w = self.w if which == 0 else self.w * np.exp(self.w.dot(self.ll) / self.w.sum() - self.ll)
        return np.bincount(flat, weights=w, minlength=stride).reshape(shape)

    def density1d_batch(self, specs, out=None, device_ptr=None, likes=False):
        from oracle.getdist_oracle import bin_indices

        assert device_ptr is None
        stride = max(s.fine_bins for s in specs)
        P, L = np.zeros((len(specs), stride)), np.zeros((len(specs), stride))
        res = []
        for i, s in enumerate(specs):
            F = s.fine_bins
            fw = (s.binmax - s.binmin) / (F - 1)
            ix = bin_indices(self.X[:, s.param], s.binmin, fw)
            bins = np.bincount(ix, weights=self.w, minlength=F)
            sp = Spec1D(*[getattr(s, f) for f, _ in Spec1D._fields_])
            r, row = Res1D(), np.empty(F)
            if likes:  # second histogram with weights * exp(mean_loglike - loglikes), as gdk_set_loglikes builds it
                lw = self.w * np.exp(self.w.dot(self.ll) / self.w.sum() - self.ll)
                lbins, lrow = np.bincount(ix, weights=lw, minlength=F), np.empty(F)
                self.hs.hs_kde1d_likes(C.byref(sp), dptr(bins), dptr(lbins), dptr(row), dptr(lrow), C.byref(r))
                L[i, :F] = lrow
            else:
                self.hs.hs_kde1d(C.byref(sp), dptr(bins), dptr(row), C.byref(r))
            P[i, :F] = row
            res.append(r)
        return (P, L, res) if likes else (P, res)


@pytest.fixture()
def fake_ctx(hs, monkeypatch):  # noqa: F811
    from getdist_b200 import _abi

    FakeContext.hs = hs
    monkeypatch.setattr(_abi, "Context", FakeContext)
    return FakeContext


@pytest.mark.parametrize("name", ["mix3", "bounded", "likes"])
def test_marge_limits_through_the_mirror(fake_ctx, name):
    """MCSamples.setMargeLimits / _setMargeLimits (batched _setDensitiesandMarge1D) against the reference's limits"""
    from getdist_b200 import MCSamples

    g = np.load(os.path.join(GOLDEN, "limits.npz"))
    case, gold = load_case(name)
    mc = MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"],
                   sampler="uncorrelated", settings=case["settings"] or None)
    np.testing.assert_allclose(mc.max_frac_twotail, g[name + "/max_frac_twotail"], rtol=1e-13)
    lims = mc.setMargeLimits()
    tags = {0: "two", 1: ">", 2: "<", 3: "none"}
    for j, nm in enumerate(case["names"]):
        ref, rt = g["%s/%d/limits" % (name, j)], g["%s/%d/tags" % (name, j)]
        assert [l.limitTag() for l in lims[nm]] == [tags[int(t)] for t in rt], (name, nm)
        got = np.array([[l.lower, l.upper] for l in lims[nm]], dtype=np.float64)
        np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7 * mc.sddev[j], err_msg=str((name, nm)))
        # ranges left by _initParam, and the cached density the limits were computed from
        par = mc.paramNames.names[j]
        gp = gold["d1/default/%d/par" % j]
        np.testing.assert_allclose([par.range_min, par.range_max, par.sigma_range], gp[:3], rtol=1e-12)
        assert np.max(np.abs(mc.get1DDensity(nm).P - gold["d1/default/%d/P" % j])) < 1e-7
    one = mc._setMargeLimits(mc.paramNames.names[0])
    assert [l.limitTag() for l in one] == [l.limitTag() for l in lims[case["names"][0]]]


def test_unknown_parameter_and_cache(fake_ctx):
    from getdist_b200 import MCSamples, ParamError

    case, _ = load_case("mix3")
    mc = MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], sampler="uncorrelated")
    assert mc.get1DDensity("nope") is None
    d = mc.get1DDensity("a")
    assert mc.get1DDensity("a") is d  # cached when called without kwargs (mcsamples.py:1669-1670)
    assert mc.get1DDensity("a", fine_bins=512) is not d
    with pytest.raises(ParamError):
        mc.setMargeLimits(["a", "zzz"])


@pytest.mark.parametrize("name", ["mix3", "unit5", "bounded", "highcorr", "periodic"])
def test_2d_planner_on_cpu(fake_ctx, name):
    """The 2D planner is pure host logic: the vectorised batch planner (_specs_2d_batch) must produce the same
    gdk_spec2d fields as the per-pair planner (_spec_2d), and both must agree with the oracle's grid geometry,
    scaled-up grid sizes (mcsamples.py:1812-1819) and branch selection on every golden case."""
    from getdist_b200 import MCSamples, _abi
    from oracle.getdist_oracle import bin_geometry

    case, g = load_case(name)
    mc = MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"],
                   sampler="uncorrelated", settings=case["settings"] or None)
    o = make_oracle(case)
    P = mc.n
    pairs = [(i, k) for i in range(P) for k in range(P) if i != k]
    mc._ensure_param_ranges(range(P))
    mc._ensure_neff(range(P))
    for kw in case["kwargs_2d"]:
        batch = mc._specs_2d_batch(pairs, kw)
        for row, (j, j2) in zip(batch, pairs):
            single = mc._spec_2d(j, j2, kw)
            for fname, _ in _abi.Spec2D._fields_:
                if fname != "contours":
                    assert row[fname] == getattr(single, fname), (name, kw, j, j2, fname)
    for (jx, jy) in case["pairs"]:
        sp = mc._spec_2d(jx, jy, {})
        d = o.density_2d(jx, jy)
        assert sp.fine_bins == d.fine_bins
        parx, pary = o.pars[jx], o.pars[jy]
        np.testing.assert_allclose([sp.xbinmin, sp.xbinmax], bin_geometry(parx, sp.fine_bins)[:2], rtol=1e-13)
        np.testing.assert_allclose([sp.ybinmin, sp.ybinmax], bin_geometry(pary, sp.fine_bins)[:2], rtol=1e-13)
        branch = d.extra.get("branch")
        if branch == "shear":
            assert sp.bw_mode == 2
        elif branch == "plain":
            assert sp.bw_mode == 1


def test_use_effective_samples_2D_is_inert(fake_ctx):
    """In the reference the setting never reaches getAutoBandwidth2D from get2DDensityGridData (its use_2D_Neff=False
    default shadows it, mcsamples.py:1297, 1327-1331 -- checked by running the reference): same specs either way."""
    from getdist_b200 import MCSamples, _abi

    case, _ = load_case("mix3")
    kw = dict(samples=case["samples"], weights=case["weights"], names=case["names"], sampler="uncorrelated")
    a, b = MCSamples(**kw), MCSamples(**kw)
    b.use_effective_samples_2D = True
    for mc in (a, b):
        mc._ensure_param_ranges(range(3))
        mc._ensure_neff(range(3))
    pairs = [(0, 1), (1, 2), (2, 0)]
    sa, sb = a._specs_2d_batch(pairs, {}), b._specs_2d_batch(pairs, {})
    assert sa.tobytes() == sb.tobytes()
    for (j, j2) in pairs:
        x, y = a._spec_2d(j, j2, {}), b._spec_2d(j, j2, {})
        assert all(getattr(x, f) == getattr(y, f) for f, _ in _abi.Spec2D._fields_ if f != "contours")


def test_fine_bins_all_matches_the_planner(fake_ctx):
    """the light fine_bins_2D rule used for the gathered multi-GPU layout == the planner's column, on strongly
    correlated data where the grids are scaled up (mcsamples.py:1811-1818)"""
    from getdist_b200 import MCSamples

    rng = np.random.default_rng(3)
    P, N = 10, 4000
    L = np.linalg.cholesky(0.95 ** np.abs(np.subtract.outer(np.arange(P), np.arange(P))))
    X = rng.standard_normal((N, P)).dot(L.T)
    mc = MCSamples(samples=X, names=["p%d" % i for i in range(P)], sampler="uncorrelated", settings={"fine_bins_2D": 256})
    idx, pairs = mc.triangle_pairs()
    mc._ensure_param_ranges(idx)
    fb = mc._fine_bins_2d_all(pairs)
    assert np.array_equal(fb, mc._specs_2d_batch(pairs, {})["fine_bins"].astype(np.int64))
    assert fb.max() > 256 and fb.min() == 256


def test_batch_record_and_overlap_helper():
    from getdist_b200.mcsamples import _BatchRecord, _overlapped

    cols = {"status": [0, 4], "hx": [0.1, 0.2], "levels": [[1.0, 2.0, 3.0, 0.0], [4.0, 5.0, 6.0, 0.0]]}
    r = _BatchRecord(cols, 1, [0.68, 0.95, 0.99])
    assert r["status"] == 4 and r["hx"] == 0.2 and r["levels"] == ([0.68, 0.95, 0.99], [4.0, 5.0, 6.0])
    assert "levels" in r and "hx" in r and "nope" not in r and r.get("nope", 7) == 7
    assert _BatchRecord(cols, 0, [])["levels"] is None
    cols["late"] = [10, 11]  # records attached after the wrapping (the library call was still in flight)
    assert r["late"] == 11
    # both sides run; results come back in order; an exception of the device side surfaces after the host work ended
    done = []
    assert _overlapped(lambda: "dev", lambda: done.append(1) or "host") == ("dev", "host") and done == [1]

    def boom():
        raise ValueError("device side failed")

    with pytest.raises(ValueError):
        _overlapped(boom, lambda: done.append(2))
    assert done == [1, 2]


def test_vectorised_param_ranges_equal_the_scalar_logic(fake_ctx):
    """MCSamples._finish_params (all parameters of a batch as array expressions) against the oracle's per-parameter
    restatement of _initParam's range / limit logic (mcsamples.py:1444-1484), bit for bit, over hard limits that are
    kept, dropped (far outside the samples) and one-sided"""
    from getdist_b200 import MCSamples
    from oracle.getdist_oracle import ParamState, finish_param_ranges

    rng = np.random.default_rng(11)
    N, P = 3000, 24
    X = rng.standard_normal((N, P)) * 10.0 ** rng.uniform(-3, 2, P) + rng.uniform(-5, 5, P)
    X[:, 3] = np.abs(X[:, 3])  # piles up at a boundary
    X[:, 4] = rng.exponential(1.0, N) ** 3  # irregular tails: the min(err, scale) branch
    w = rng.exponential(1.0, N)
    names = ["p%d" % i for i in range(P)]
    ranges = {}
    for i in range(P):
        lo, hi = X[:, i].min(), X[:, i].max()
        d = hi - lo
        kind = i % 6
        if kind == 1:
            ranges[names[i]] = (lo - 1e-3 * d, None)  # kept
        elif kind == 2:
            ranges[names[i]] = (lo - 5 * d, hi + 5 * d)  # both dropped
        elif kind == 3:
            ranges[names[i]] = (None, hi)  # kept, at the sample maximum
        elif kind == 4:
            ranges[names[i]] = (lo - 1e-3 * d, hi + 5 * d)  # one kept, one dropped
    mc = MCSamples(samples=X, weights=w, names=names, ranges=ranges, sampler="uncorrelated")
    fr = mc._range_fracs()
    table = mc._ctx.weighted_quantiles(list(range(P)), fr)
    mc._finish_params(list(range(P)), table)
    kept = 0
    for j, par in enumerate(mc.paramNames.names):
        ref = ParamState(name=par.name)
        ref.limmin, ref.limmax = mc.ranges.getLower(par.name), mc.ranges.getUpper(par.name)
        ref.err, ref.mean, ref.param_min, ref.param_max = mc.sddev[j], mc.means[j], X[:, j].min(), X[:, j].max()
        finish_param_ranges(ref, table[j])
        for a in ("range_min", "range_max", "sigma_range", "has_limits_bot", "has_limits_top", "has_limits", "param_min", "param_max", "err"):
            assert getattr(par, a) == getattr(ref, a), (j, a, getattr(par, a), getattr(ref, a))
        kept += par.has_limits
    assert 0 < kept < P
    # one parameter through the single-parameter entry point: same numbers
    one = MCSamples(samples=X, weights=w, names=names, ranges=ranges, sampler="uncorrelated")
    p5 = one._initParamRanges(4)
    assert p5.range_min == mc.paramNames.names[4].range_min and p5.sigma_range == mc.paramNames.names[4].sigma_range
