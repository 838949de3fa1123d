"""The shipped drop-in adapter (getdist_b200/reference_backend.py: a subclass of the reference's MCSamples) against the
unmodified reference, on the CPU: the ctypes Context is replaced by the test double of test_host_mirror_cpu.py (numpy
moments / exact quantiles, hostsim 1D grids) extended with a 2D stand-in that answers gdk_density2d_batch from the
oracle.  What is checked here is the ADAPTER -- routing, object types, ParamInfo side effects, Gelman-Rubin with the
chain boundaries in place before the first upload, re-synchronisation of settings and ranges, limits, pickling -- the
device arithmetic is checked on the B200 (tests/test_gpu_reference_backend.py)."""
import copy
import pickle

import numpy as np
import pytest

from helpers import load_case
from test_host_mirror_cpu import FakeContext
from test_hostsim import hs  # noqa: F401


class FakeContext2D(FakeContext):
    """adds gdk_density2d_batch: grids from the oracle object the test registers, contour levels by the host routine"""

    oracle = None

    def density2d_batch(self, specs, out=None, device_ptr=None, likes=False):
        from getdist_b200 import _abi
        from getdist_b200.densities import getContourLevels

        assert not likes and device_ptr is None
        fb = specs["fine_bins"].astype(np.int64)
        offsets = np.zeros(len(specs), dtype=np.int64)
        offsets[1:] = np.cumsum(fb * fb)[:-1]
        buf = np.empty(int((fb * fb).sum())) if out is None else out  # large batches hand in the buffer they wrap
        res = []
        for sp, off in zip(specs, offsets):
            d = type(self).oracle.density_2d(int(sp["px"]), int(sp["py"]))
            G = int(sp["fine_bins"])
            assert d.P.shape == (G, G)
            buf[off: off + G * G] = d.P.ravel()
            r = _abi.Result2D()
            nc = int(sp["n_contours"])
            if nc:
                lv = getContourLevels(d.P, [float(c) for c in sp["contours"][:nc]])
                for k in range(nc):
                    r.levels[k] = lv[k]
            res.append(r)
        return buf, offsets, res


@pytest.fixture()
def backend(hs, monkeypatch, getdist_ref):  # noqa: F811
    from getdist_b200 import _abi

    FakeContext2D.hs = hs
    monkeypatch.setattr(_abi, "Context", FakeContext2D)
    from getdist_b200 import reference_backend

    return reference_backend


def _kw(case):
    return dict(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"] or None,
                sampler=case.get("sampler", "uncorrelated"), loglikes=case.get("loglikes"))


def test_moments_and_1d(backend, getdist_ref):
    case, _ = load_case("bounded")
    ref = getdist_ref.MCSamples(**_kw(case))
    mc = backend.MCSamples(**_kw(case))
    assert isinstance(mc, getdist_ref.MCSamples)
    np.testing.assert_allclose(mc.getMeans(), ref.getMeans(), rtol=1e-12)
    np.testing.assert_allclose(mc.getVars(), ref.getVars(), rtol=1e-11)
    np.testing.assert_allclose(mc.getCov(), ref.getCov(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(mc.getCorrelationMatrix(), ref.getCorrelationMatrix(), rtol=1e-10, atol=1e-13)
    assert mc.mean_mult == ref.mean_mult and mc.max_mult == ref.max_mult
    for name in case["names"][:3]:
        a, b = mc.get1DDensityGridData(name), ref.get1DDensityGridData(name)
        assert type(a) is type(b)  # the reference's own Density1D
        assert np.max(np.abs(a.P - b.P)) < 1e-7
        np.testing.assert_allclose(a.view_ranges, b.view_ranges, rtol=1e-12)
        pa, pb = mc.paramNames.parWithName(name), ref.paramNames.parWithName(name)
        for attr in ("err", "mean", "range_min", "range_max", "sigma_range", "param_min", "param_max"):
            np.testing.assert_allclose(getattr(pa, attr), getattr(pb, attr), rtol=1e-12)
        assert (pa.has_limits_bot, pa.has_limits_top) == (pb.has_limits_bot, pb.has_limits_top)
        assert mc.get1DDensity(name) is mc.density1D[name]
    fr = np.array([0.025, 0.975])
    assert np.array_equal(mc.confidence(1, fr), ref.confidence(1, fr))
    assert np.array_equal(mc.confidence(1, fr, start=10, end=3000), ref.confidence(1, fr, start=10, end=3000))
    assert np.array_equal(mc.confidence(mc.samples[:, 1] * 2, fr), ref.confidence(ref.samples[:, 1] * 2, fr))  # inherited


def test_gelman_rubin_and_chainlist(backend, getdist_ref):
    """chain boundaries reach the device before the first upload (the round-1 stub divided by nchains - 1 = 0)"""
    case, g = load_case("chains")
    kw = _kw(case)
    mc = backend.MCSamples(**kw)
    ref = getdist_ref.MCSamples(**kw)
    np.testing.assert_allclose(mc.getGelmanRubin(), float(g["gelman_rubin"]), rtol=1e-9)
    np.testing.assert_allclose(mc.getGelmanRubin(3), ref.getGelmanRubin(3), rtol=1e-9)
    np.testing.assert_allclose(mc.getGelmanRubinEigenvalues(), ref.getGelmanRubinEigenvalues(), rtol=1e-7, atol=1e-12)
    chains = ref.getSeparateChains()[:3]  # explicit chainlist of reference objects: routed to the inherited method
    np.testing.assert_allclose(mc.getGelmanRubin(chainlist=chains), ref.getGelmanRubin(chainlist=chains), rtol=1e-12)


def test_2d_contours_settings_resync_and_limits(backend, getdist_ref):
    from helpers import make_oracle

    case, _ = load_case("mix3")
    ref = getdist_ref.MCSamples(**_kw(case))
    mc = backend.MCSamples(**_kw(case))
    FakeContext2D.oracle = make_oracle(case)
    d = mc.get2DDensityGridData("a", "b", num_plot_contours=2)  # what plots.py:641 asks for
    r = ref.get2DDensityGridData("a", "b", num_plot_contours=2)
    assert type(d) is type(r) and d.P.shape == r.P.shape
    assert np.max(np.abs(d.P - r.P)) < 1e-9
    np.testing.assert_allclose(d.contours, r.contours, rtol=1e-8)
    assert d.likes is None and len(d.contours) == 2
    np.testing.assert_allclose(np.asarray(d.view_ranges), np.asarray(r.view_ranges), rtol=1e-12)
    dd = mc.get2DDensity("a", "b", normalized=True)  # normalises its own copy; the next call is max-normalised again
    assert abs(mc.get2DDensityGridData("a", "b", get_density=True).P.max() - 1) < 1e-15 and dd.P.max() != 1
    assert mc.get2DDensityGridData("a", "nope") is None
    # a later updateSettings on the reference-facing object reaches the device planner
    g = mc._gpu()
    mc.updateSettings({"fine_bins": 512, "smooth_scale_1D": 0.4})
    ref.updateSettings({"fine_bins": 512, "smooth_scale_1D": 0.4})
    assert mc._gpu() is g and g.fine_bins == 512 and g.smooth_scale_1D == 0.4 and not g.density1D
    a, b = mc.get1DDensityGridData("c"), ref.get1DDensityGridData("c")
    assert a.P.size == 512 and np.max(np.abs(a.P - b.P)) < 1e-9
    # hard ranges set afterwards (setRanges) reach it as well
    mc.setRanges({"b": (2.6, None)})
    ref.setRanges({"b": (2.6, None)})
    mc.updateBaseStatistics()
    ref.updateBaseStatistics()
    a, b = mc.get1DDensityGridData("b"), ref.get1DDensityGridData("b")
    np.testing.assert_allclose(a.view_ranges, b.view_ranges, rtol=1e-12)
    assert mc.paramNames.parWithName("b").has_limits_bot == ref.paramNames.parWithName("b").has_limits_bot
    assert np.max(np.abs(a.P - b.P)) < 1e-7
    # marginalised limits through getMargeStats (reference: testLimits, getdist_test.py:128-142)
    ms, mr = mc.getMargeStats(), ref.getMargeStats()
    for name in case["names"]:
        for k in range(2):
            la, lb = ms.parWithName(name).limits[k], mr.parWithName(name).limits[k]
            assert la.limitTag() == lb.limitTag()
            np.testing.assert_allclose([la.lower, la.upper], [lb.lower, lb.upper], rtol=1e-6)


def test_pickle_and_copy(backend, getdist_ref):
    case, _ = load_case("mix3")
    mc = backend.MCSamples(**_kw(case))
    m0 = mc.getMeans().copy()
    mc._gpu()
    for k, other in enumerate((pickle.loads(pickle.dumps(mc)), copy.deepcopy(mc), mc.copy(settings={"fine_bins": 256}))):
        assert type(other) is type(mc) and other.__dict__.get("_gpu_obj") is not mc._gpu()
        assert k == 2 or "_gpu_obj" not in other.__dict__  # the device object is not part of the state: rebuilt lazily
        np.testing.assert_allclose(other.getMeans(), m0, rtol=1e-13)
        assert other.get1DDensityGridData("a").P.size == other.fine_bins
    # new weights invalidate the device copy (chains.py:310-323)
    g = mc._gpu()
    mc.reweightAddingLogLikes(np.zeros(mc.numrows)) if mc.loglikes is not None else mc.setSamples(mc.samples, mc.weights * 2.0)
    assert mc._gpu() is not g


def test_large_batch_is_wrapped_while_the_call_is_in_flight(hs, monkeypatch):  # noqa: F811
    """>= 64 pairs: the mirror allocates the result buffer, wraps views of it in the calling thread while the library
    call runs in a worker thread, and attaches the result records afterwards (mcsamples._overlapped)"""
    from getdist_b200 import MCSamples, _abi
    from oracle.getdist_oracle import OracleSamples

    FakeContext2D.hs = hs
    monkeypatch.setattr(_abi, "Context", FakeContext2D)
    rng = np.random.default_rng(8)
    P, N = 12, 2500
    X = rng.standard_normal((N, P)) * np.arange(1, P + 1)
    names = ["p%d" % i for i in range(P)]
    FakeContext2D.oracle = OracleSamples(X, None, names=names, sampler="uncorrelated", settings={"fine_bins_2D": 64})
    mc = MCSamples(samples=X, names=names, sampler="uncorrelated", settings={"fine_bins_2D": 64})
    idx, pairs = mc.triangle_pairs()
    assert len(pairs) >= 64
    d2 = mc._densities_2d(pairs)
    for (a, b), d in zip(pairs[::7], d2[::7]):
        ref = FakeContext2D.oracle.density_2d(a, b)
        assert np.array_equal(d.P, ref.P) and not d.P.flags.writeable
        assert d._gdk["status"] == 0 and len(d._gdk["levels"][1]) == 3
        assert mc._density2D[(a, b)] is d
    fresh = mc.get2DDensityGridData(pairs[3][0], pairs[3][1], num_plot_contours=2)
    assert fresh.P.flags.writeable and len(fresh.contours) == 2
