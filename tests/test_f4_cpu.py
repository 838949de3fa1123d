"""SURVEY s8f-4 on the CPU: (1) the oracle's restatement of getRawNDDensityGridData and of the MeanVar / split
convergence tests against goldens produced by the unmodified reference (tests/golden/f4.npz, make_golden.py f4);
(2) the host mirror's getRawNDDensityGridData / getConvergeTests / getFractionIndices through the test-double context
against the same goldens (the reference's own text for getConvergeTests)."""
import os

import numpy as np
import pytest

from cases import CASES, input_digest
from helpers import GOLDEN, make_oracle
from test_host_mirror_cpu import fake_ctx  # noqa: F401
from test_hostsim import hs  # noqa: F401

F4 = ("chains", "likes", "bounded")
ND_SETS = ([0, 1, 2], [1, 0], [0, 1, 2, 3])


def _norm_text(t):
    return t.replace("-0.00000", " 0.00000")  # an eigenvalue that is zero to rounding may print either sign


def _g():
    return np.load(os.path.join(GOLDEN, "f4.npz"))


def _case(name):
    case = CASES[name]()
    assert str(_g()[name + "/digest"]) == input_digest(case)
    return case


@pytest.mark.parametrize("name", F4)
def test_oracle_raw_nd_and_converge(name):
    g = _g()
    case = _case(name)
    case = dict(case, meanlikes=case.get("loglikes") is not None)
    o = make_oracle(case)
    likes = o.loglikes is not None
    for js in ND_SETS:
        js = js[: o.n]
        tag = "_".join(str(j) for j in js)
        xs, P, L, M = o.raw_nd_density(js, meanlikes=likes, maxlikes=likes)
        np.testing.assert_allclose(P, g["%s/nd/%s/P" % (name, tag)], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(xs[0], g["%s/nd/%s/x0" % (name, tag)], rtol=1e-14)
        if likes:
            np.testing.assert_allclose(L, g["%s/nd/%s/likes" % (name, tag)], rtol=1e-11, atol=1e-15)
            np.testing.assert_allclose(M, g["%s/nd/%s/maxlikes" % (name, tag)], rtol=1e-13, atol=0)
    for n in (2, 3, 4):
        assert np.array_equal(o.fraction_indices(n), g["%s/frac/%d" % (name, n)])
    np.testing.assert_allclose(o.split_tests(), g[name + "/split_tests"], rtol=1e-12, atol=1e-15)


def _mirror(case):
    from getdist_b200 import MCSamples

    return MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"],
                     sampler=case.get("sampler", "uncorrelated"), settings=case["settings"] or None, loglikes=case.get("loglikes"))


@pytest.mark.parametrize("name", F4)
def test_mirror_raw_nd_and_converge(fake_ctx, name):  # noqa: F811
    g = _g()
    case = _case(name)
    mc = _mirror(case)
    likes = mc.loglikes is not None
    for js in ND_SETS:
        js = js[: mc.n]
        tag = "_".join(str(j) for j in js)
        d = mc.getRawNDDensityGridData(js, meanlikes=likes, maxlikes=likes)
        np.testing.assert_allclose(d.P, g["%s/nd/%s/P" % (name, tag)], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(d.contours, g["%s/nd/%s/contours" % (name, tag)], rtol=1e-10)
        np.testing.assert_allclose(d.xs[0], g["%s/nd/%s/x0" % (name, tag)], rtol=1e-14)
        assert d.P.shape == tuple([mc.num_bins_ND] * len(js))
        if likes:
            np.testing.assert_allclose(d.likes, g["%s/nd/%s/likes" % (name, tag)], rtol=1e-11, atol=1e-15)
            np.testing.assert_allclose(d.maxlikes, g["%s/nd/%s/maxlikes" % (name, tag)], rtol=1e-13)
            np.testing.assert_allclose(d.maxcontours, g["%s/nd/%s/maxcontours" % (name, tag)], rtol=1e-10)
    for n in (2, 3, 4):
        assert np.array_equal(mc.getFractionIndices(mc.weights, n), g["%s/frac/%d" % (name, n)])
    np.testing.assert_allclose(mc.getSplitTests(), g[name + "/split_tests"], rtol=1e-12, atol=1e-15)
    if mc.chain_offsets is not None:
        assert _norm_text(mc.getConvergeTests()) == _norm_text(str(g[name + "/converge_text"]))  # the reference's own text, character by character
        # chainlist: a subset of the chains, by ChainView or by index
        ch = mc.getSeparateChains()
        a = mc.getGelmanRubin(chainlist=ch[:3])
        b = mc.getGelmanRubin(chainlist=[0, 1, 2])
        assert a == b and a != mc.getGelmanRubin()


def test_mirror_confidence_forms(fake_ctx):  # noqa: F811
    """confidence() on stored columns with a row range, on arbitrary vectors and with other weights (chains.py:793-838)"""
    from oracle.getdist_oracle import weighted_quantiles

    case = _case("bounded")
    mc = _mirror(case)
    X, w = case["samples"], case["weights"]
    fr = np.array([0.025, 0.5, 0.975])
    assert np.array_equal(mc.confidence(1, fr, start=100, end=20000), weighted_quantiles(X[100:20000, 1], w[100:20000], fr))
    v = X[:, 0] * 2 + X[:, 1]
    assert np.array_equal(mc.confidence(v, fr), weighted_quantiles(v, w, fr))
    w2 = np.sqrt(w)
    assert np.array_equal(mc.confidence(X[:, 2], fr, weights=w2, upper=True), weighted_quantiles(X[:, 2], w2, 1 - fr))
    assert mc.confidence(v, 0.3) == weighted_quantiles(v, w, np.array([0.3]))[0]
