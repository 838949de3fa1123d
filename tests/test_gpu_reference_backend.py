"""The drop-in adapter (getdist_b200.reference_backend.MCSamples, a subclass of the reference class) on the B200,
side by side with the UNMODIFIED reference running on the box's CPU (imported from baseline/_ref, the offline install
of /root/reference that travels with the snapshot; skipped where it is absent).  Same inputs, same calls as
getdist.plots issues them (plots.py:616, 641, 1116): densities within the 1e-6 bar, contours, limits, Gelman-Rubin."""
import pickle

import numpy as np
import pytest

from helpers import load_case

pytestmark = pytest.mark.gpu


def _kw(case):
    return dict(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"] or None,
                sampler=case.get("sampler", "uncorrelated"), loglikes=case.get("loglikes"), settings=case["settings"] or None)


@pytest.fixture(scope="module")
def both(getdist_ref):
    from getdist_b200 import reference_backend

    def make(name):
        case, g = load_case(name)
        return case, g, reference_backend.MCSamples(**_kw(case)), getdist_ref.MCSamples(**_kw(case))

    return make


@pytest.mark.parametrize("name", ["mix3", "bounded", "chains", "likes"])
def test_adapter_matches_reference(both, name):
    case, g, mc, ref = both(name)
    np.testing.assert_allclose(mc.getMeans(), ref.getMeans(), rtol=1e-12, atol=1e-12 * np.sqrt(ref.getVars()).max())
    scale = np.sqrt(np.outer(np.diag(ref.getCov()), np.diag(ref.getCov())))
    assert np.max(np.abs(mc.getCov() - ref.getCov()) / scale) < 1e-12
    if ref.chain_offsets is not None:
        np.testing.assert_allclose(mc.getGelmanRubin(), ref.getGelmanRubin(), rtol=1e-9)
    for j, nm in enumerate(case["names"]):
        a, b = mc.get1DDensityGridData(nm), ref.get1DDensityGridData(nm)
        assert type(a) is type(b) and np.max(np.abs(a.P - b.P)) < 1e-6, (nm, np.max(np.abs(a.P - b.P)))
        np.testing.assert_allclose(a.view_ranges, b.view_ranges, rtol=1e-12)
        fr = np.array([0.05, 0.95])  # the reference multiplies limfrac by the norm: an array, not a list (chains.py:833)
        assert np.array_equal(mc.confidence(j, fr), ref.confidence(j, fr))
    likes = ref.loglikes is not None and name == "likes"
    for (jx, jy) in case["pairs"][:3]:
        a = mc.get2DDensityGridData(jx, jy, num_plot_contours=2, meanlikes=likes)
        b = ref.get2DDensityGridData(jx, jy, num_plot_contours=2, meanlikes=likes)
        assert type(a) is type(b)
        tnc = bool(mc._gpu()._density2D.get((jx, jy)) and mc._gpu()._density2D[(jx, jy)]._gdk["status"] & (64 | 128))
        assert np.max(np.abs(a.P - b.P)) < (1e-5 if tnc or likes else 1e-6), (jx, jy, np.max(np.abs(a.P - b.P)))
        np.testing.assert_allclose(a.contours, b.contours, rtol=2e-4 if tnc else 1e-5)
        if likes:
            assert np.max(np.abs(a.likes - b.likes)) < 1e-5
    ms, mr = mc.getMargeStats(), ref.getMargeStats()
    for nm in case["names"]:
        for k in range(2):
            la, lb = ms.parWithName(nm).limits[k], mr.parWithName(nm).limits[k]
            assert la.limitTag() == lb.limitTag()
            np.testing.assert_allclose([la.lower, la.upper], [lb.lower, lb.upper], rtol=1e-6)
    d, r = mc.getRawNDDensityGridData(case["names"][:3]), ref.getRawNDDensityGridData(case["names"][:3])
    assert type(d) is type(r)
    np.testing.assert_allclose(d.P, r.P, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(d.contours, r.contours, rtol=1e-8)


def test_adapter_mask_function_settings_and_pickle(both):
    from cases import CASES

    case, g, mc, ref = both("likes")
    mf = CASES["likes"]()["mask_function"]
    jx, jy = case["mask_pairs"][0]
    a = mc.get2DDensityGridData(jx, jy, mask_function=mf, get_density=True)
    b = ref.get2DDensityGridData(jx, jy, mask_function=mf, get_density=True)
    assert np.array_equal(np.asarray(a.mask), np.asarray(b.mask))
    ok = ~np.asarray(b.mask)
    assert np.max(np.abs(a.P[ok] - b.P[ok])) < 1e-5
    mc.updateSettings({"fine_bins_2D": 128, "smooth_scale_2D": 0.5})
    ref.updateSettings({"fine_bins_2D": 128, "smooth_scale_2D": 0.5})
    a, b = mc.get2DDensity(0, 1), ref.get2DDensity(0, 1)
    assert a.P.shape == (128, 128) and np.max(np.abs(a.P - b.P)) < 1e-9
    other = pickle.loads(pickle.dumps(mc))
    assert np.max(np.abs(other.get2DDensity(0, 1).P - b.P)) < 1e-9
