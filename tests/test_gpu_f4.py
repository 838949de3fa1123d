"""GPU parity of SURVEY s8f-4 (through the C-ABI): raw ND densities (gdk_histnd), fraction indices
(gdk_weight_fraction_rows), split / MeanVar convergence tests (gdk_weighted_quantiles_range + per-chain moments) against
goldens produced by the unmodified reference (tests/golden/f4.npz), and the transparent batching of a serial caller."""
import os

import numpy as np
import pytest

from cases import CASES, input_digest
from helpers import GOLDEN, load_case

pytestmark = pytest.mark.gpu

F4 = ("chains", "likes", "bounded")
ND_SETS = ([0, 1, 2], [1, 0], [0, 1, 2, 3])


def _norm_text(t):
    return t.replace("-0.00000", " 0.00000")  # an eigenvalue that is zero to rounding may print either sign


def _gpu(case, **kw):
    from getdist_b200 import MCSamples

    return MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"],
                     sampler=case.get("sampler", "uncorrelated"), settings=case["settings"] or None, loglikes=case.get("loglikes"), **kw)


@pytest.mark.parametrize("name", F4)
def test_raw_nd_and_converge(name):
    g = np.load(os.path.join(GOLDEN, "f4.npz"))
    case = CASES[name]()
    assert str(g[name + "/digest"]) == input_digest(case)
    mc = _gpu(case)
    likes = mc.loglikes is not None
    for js in ND_SETS:
        js = js[: mc.n]
        tag = "_".join(str(j) for j in js)
        d = mc.getRawNDDensityGridData(js, meanlikes=likes, maxlikes=likes)
        np.testing.assert_allclose(d.P, g["%s/nd/%s/P" % (name, tag)], rtol=1e-11, atol=1e-15)  # fixed-point weights: 2^-37
        np.testing.assert_allclose(d.contours, g["%s/nd/%s/contours" % (name, tag)], rtol=1e-9)
        if likes:
            np.testing.assert_allclose(d.likes, g["%s/nd/%s/likes" % (name, tag)], rtol=1e-10, atol=1e-15)
            np.testing.assert_allclose(d.maxlikes, g["%s/nd/%s/maxlikes" % (name, tag)], rtol=1e-12)
            np.testing.assert_allclose(d.maxcontours, g["%s/nd/%s/maxcontours" % (name, tag)], rtol=1e-9)
    for n in (2, 3, 4):
        assert np.array_equal(mc.getFractionIndices(mc.weights, n), g["%s/frac/%d" % (name, n)])
    np.testing.assert_allclose(mc.getSplitTests(), g[name + "/split_tests"], rtol=1e-11, atol=1e-14)
    if mc.chain_offsets is not None:
        assert _norm_text(mc.getConvergeTests()) == _norm_text(str(g[name + "/converge_text"]))


def test_confidence_forms():
    from oracle.getdist_oracle import weighted_quantiles

    case, _ = load_case("bounded")
    mc = _gpu(case)
    X, w = case["samples"], case["weights"]
    fr = np.array([0.025, 0.5, 0.975])
    assert np.array_equal(mc.confidence(1, fr, start=101, end=20001), weighted_quantiles(X[101:20001, 1], w[101:20001], fr))
    v = X[:, 0] * 2 + X[:, 1]
    assert np.array_equal(mc.confidence(v, fr), weighted_quantiles(v, w, fr))
    w2 = np.sqrt(w)
    assert np.array_equal(mc.confidence(X[:, 2], fr, weights=w2, upper=True), weighted_quantiles(X[:, 2], w2, 1 - fr))


def test_serial_caller_is_batched():
    """A caller that asks pair by pair in the order of getdist.plots.triangle_plot (plots.py:2613) gets batched launches
    without calling prefetch_triangle: the first miss computes the whole triangle (few parameters) or one plot row at a
    time (many parameters), and the results equal the single calls."""
    case, _ = load_case("unit5")
    ref = _gpu(case)
    ref.auto_prefetch = False
    for limit in (40, 2):  # whole triangle at once / row batching
        mc = _gpu(case)
        mc.auto_prefetch_max_params = limit
        calls = []
        orig = mc._ctx.density2d_batch
        mc._ctx.density2d_batch = lambda specs, **kw: (calls.append(len(specs)), orig(specs, **kw))[1]
        n = mc.n
        for i in range(n):
            mc.get1DDensity(i)
            for i2 in range(i):
                d = mc.get2DDensityGridData(i2, i, num_plot_contours=2)
                r = ref.get2DDensityGridData(i2, i, num_plot_contours=2)
                assert np.array_equal(d.P, r.P) and np.allclose(d.contours, r.contours, rtol=1e-12)
        assert len(calls) == (1 if limit == 40 else n - 1), calls
        # the cache keeps a private copy: normalising a returned grid in place leaves later calls untouched
        a = mc.get2DDensity(0, 1, normalized=True)
        b = mc.get2DDensityGridData(0, 1, get_density=True)
        assert abs(b.P.max() - 1.0) < 1e-15 and a.P.max() != 1.0
