"""Generate golden vectors by running the UNMODIFIED reference (getdist imported from
/root/reference) on the seeded cases in cases.py.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs tests/golden/<case>.npz.  The reference's private bandwidth helpers are wrapped (not
modified) so that the bandwidths they return can be recorded next to the density grids.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import getdist  # noqa: E402
from getdist import MCSamples  # noqa: E402

from cases import CASES, grid_stride, input_digest, kw_tag  # noqa: E402


def make_c1_inputs():
    """C1 samples from the reference's own generators -> c1_inputs.npz (cases.py:_c1_inputs)"""
    from getdist.gaussian_mixtures import MixtureND, RandomTestMixtureND

    names = ["x", "y", "z"]
    rand = RandomTestMixtureND(ndim=3, ncomponent=2, seed=10, names=names).MCSamples(100000, random_state=10).samples
    s = 2.0 / 3
    cov = np.diag([s * s, s * s, 1.5 ** 2])
    wj = MixtureND([[-1.0, 0.0, 0.5], [1.0, 0.0, 0.5]], [cov, cov], names=names).MCSamples(100000, random_state=10).samples
    w = np.random.default_rng(11).exponential(1.0, 100000)
    np.savez_compressed(os.path.join(HERE, "c1_inputs.npz"), rand=np.ascontiguousarray(rand), wj=np.ascontiguousarray(wj), w=w)
    print("c1 inputs ok", rand.shape, wj.shape)


def run_case(name):
    case = CASES[name]()
    out = {"digest": np.array(input_digest(case)), "getdist_version": np.array(getdist.__version__)}
    mc = MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"],
                   ranges=case["ranges"] or None, sampler=case.get("sampler", "uncorrelated"), loglikes=case.get("loglikes"), settings=case["settings"] or None)
    rec = {}
    orig1d = mc.getAutoBandwidth1D
    orig2d = mc.getAutoBandwidth2D

    def wrap1d(*a, **k):
        r = orig1d(*a, **k)
        rec["h1d"] = r
        return r

    def wrap2d(*a, **k):
        r = orig2d(*a, **k)
        rec["h2d"] = r
        return r

    mc.getAutoBandwidth1D = wrap1d
    mc.getAutoBandwidth2D = wrap2d

    out["means"] = mc.getMeans()
    out["vars"] = mc.getVars()
    out["cov"] = mc.getCov()
    out["corr"] = mc.getCorrelationMatrix()
    out["norm"] = np.array(mc.norm)
    out["max_mult"] = np.array(mc.max_mult)
    out["mean_mult"] = np.array(mc.mean_mult)
    if mc.chain_offsets is not None:
        out["chain_offsets"] = np.asarray(mc.chain_offsets)
        out["gelman_rubin"] = np.array(mc.getGelmanRubin())
        out["gelman_rubin_eig"] = mc.getGelmanRubinEigenvalues()
        out["gelman_rubin_3"] = np.array(mc.getGelmanRubin(nparam=3))
    P = len(case["names"])
    # quantiles used by _initParam, straight from confidence()
    fr = np.array([0.001, 0.999] + list(np.linspace(0.1, 0.9, 9)))
    out["quantile_fracs"] = fr
    out["quantiles"] = np.array([mc.confidence(j, fr) for j in range(P)])
    for kw in case["kwargs_1d"]:
        tag = kw_tag(kw)
        for j in range(P):
            rec.clear()
            d = mc.get1DDensityGridData(j, **kw)
            par = mc.paramNames.names[j]
            out["d1/%s/%d/P" % (tag, j)] = d.P
            out["d1/%s/%d/x" % (tag, j)] = np.array([d.x[0], d.x[-1], d.x.size])
            out["d1/%s/%d/par" % (tag, j)] = np.array(
                [par.range_min, par.range_max, par.sigma_range, par.param_min, par.param_max, par.err, par.mean,
                 float(par.has_limits_bot), float(par.has_limits_top),
                 par.kde_h if getattr(par, "kde_h", None) is not None else np.nan,
                 rec.get("h1d", np.nan), par.N_eff_kde if getattr(par, "N_eff_kde", None) is not None else np.nan])
    for kw in case["kwargs_2d"]:
        tag = kw_tag(kw)
        for (jx, jy) in case["pairs"]:
            rec.clear()
            d = mc.get2DDensity(jx, jy, **kw)
            st = grid_stride(d.P.shape[0])
            out["d2/%s/%d_%d/P" % (tag, jx, jy)] = d.P[::st, ::st]
            out["d2/%s/%d_%d/xy" % (tag, jx, jy)] = np.array([d.x[0], d.x[-1], d.x.size, d.y[0], d.y[-1], d.y.size])
            out["d2/%s/%d_%d/h" % (tag, jx, jy)] = np.array(rec.get("h2d", (np.nan, np.nan, np.nan)), dtype=np.float64)
            if not kw:
                dc = mc.get2DDensityGridData(jx, jy, num_plot_contours=3)
                out["d2/%s/%d_%d/contours" % (tag, jx, jy)] = np.asarray(dc.contours)
    if case.get("meanlikes"):
        out["mean_loglike"] = np.array(mc.mean_loglike)
        for kw in case["likes_kwargs_1d"]:
            tag = kw_tag(kw)
            for j in range(P):
                d = mc.get1DDensityGridData(j, meanlikes=True, **kw)
                out["l1/%s/%d/P" % (tag, j)] = d.P
                out["l1/%s/%d/likes" % (tag, j)] = d.likes
        for kw in case["likes_kwargs_2d"]:
            tag = kw_tag(kw)
            for (jx, jy) in case["pairs"]:
                d = mc.get2DDensityGridData(jx, jy, meanlikes=True, **kw)
                out["l2/%s/%d_%d/P" % (tag, jx, jy)] = d.P
                out["l2/%s/%d_%d/likes" % (tag, jx, jy)] = d.likes
                out["l2/%s/%d_%d/contours" % (tag, jx, jy)] = np.asarray(d.contours)
    if case.get("mask_function"):
        for kw in case["mask_kwargs_2d"]:
            tag = kw_tag(kw)
            for (jx, jy) in case["mask_pairs"]:
                d = mc.get2DDensityGridData(jx, jy, mask_function=case["mask_function"], get_density=True, **kw)
                out["m2/%s/%d_%d/P" % (tag, jx, jy)] = d.P
                out["m2/%s/%d_%d/mask" % (tag, jx, jy)] = np.asarray(d.mask)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", len(out), "arrays")


LIMIT_CASES = ("mix3", "bounded", "likes")


def run_limits():
    """Marginalised limits of the reference (MCSamples._setMargeLimits, mcsamples.py:2460-2531) for every parameter of
    a few cases -> limits.npz: per parameter the (lower, upper) pairs of the three default contours and the limit tags
    encoded as 0 two, 1 '>', 2 '<', 3 none."""
    code = {"two": 0, ">": 1, "<": 2, "none": 3}
    out = {}
    for name in LIMIT_CASES:
        case = CASES[name]()
        mc = MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"] or None,
                       sampler=case.get("sampler", "uncorrelated"), loglikes=case.get("loglikes"), settings=case["settings"] or None)
        out[name + "/digest"] = np.array(input_digest(case))
        out[name + "/contours"] = np.asarray(mc.contours)
        out[name + "/max_frac_twotail"] = np.asarray(mc.max_frac_twotail)
        for j, par in enumerate(mc.paramNames.names):
            conf = mc.initParamConfidenceData(mc.samples[:, j])
            mc._setMargeLimits(par, conf)
            out["%s/%d/limits" % (name, j)] = np.array([[np.nan if l.lower is None else l.lower,
                                                         np.nan if l.upper is None else l.upper] for l in par.limits])
            out["%s/%d/tags" % (name, j)] = np.array([code[l.limitTag()] for l in par.limits])
    np.savez_compressed(os.path.join(HERE, "limits.npz"), **out)
    print("limits ok", len(out), "arrays")


F4_CASES = ("chains", "likes", "bounded")


def run_f4():
    """SURVEY s8f-4 outputs of the unmodified reference -> f4.npz: raw ND densities (getRawNDDensityGridData,
    mcsamples.py:2098-2235) incl. mean / profile likelihoods, the fraction indices, and the MeanVar / GelmanRubin /
    SplitTest sections of getConvergeTests (mcsamples.py:905-1034) as text plus the numbers behind them."""
    out = {}
    for name in F4_CASES:
        case = CASES[name]()
        mc = MCSamples(samples=case["samples"], weights=case["weights"], names=case["names"], ranges=case["ranges"] or None,
                       sampler=case.get("sampler", "uncorrelated"), loglikes=case.get("loglikes"), settings=case["settings"] or None)
        out[name + "/digest"] = np.array(input_digest(case))
        likes = case.get("loglikes") is not None
        for js in ([0, 1, 2], [1, 0], [0, 1, 2, 3][: len(case["names"])]):
            tag = "_".join(str(j) for j in js)
            d = mc.getRawNDDensityGridData(js, meanlikes=likes, maxlikes=likes)
            out["%s/nd/%s/P" % (name, tag)] = d.P
            out["%s/nd/%s/contours" % (name, tag)] = np.asarray(d.contours)
            out["%s/nd/%s/x0" % (name, tag)] = np.asarray(d.xs[0])
            if likes:
                out["%s/nd/%s/likes" % (name, tag)] = d.likes
                out["%s/nd/%s/maxlikes" % (name, tag)] = d.maxlikes
                out["%s/nd/%s/maxcontours" % (name, tag)] = np.asarray(d.maxcontours)
        for n in (2, 3, 4):
            out["%s/frac/%d" % (name, n)] = mc.getFractionIndices(mc.weights, n)
        if mc.chain_offsets is not None:  # getConvergeTests starts with getSeparateChains(): combined chains only
            if mc.loglikes is None:  # getSeparateChains slices loglikes (chains.py:1519-1525)
                mc.loglikes = np.zeros(mc.numrows)
            out[name + "/converge_text"] = np.array(mc.getConvergeTests(what=("MeanVar", "GelmanRubin", "SplitTest")))
        # the split-test numbers from the reference's own confidence(..., start, end) calls (mcsamples.py:1013-1029)
        limits = np.array([1 - (1 - 0.95) / 2, (1 - 0.95) / 2])
        st = np.zeros((mc.n, mc.max_split_tests - 1, 2))
        for j in range(mc.n):
            confids = mc.confidence(mc.samples[:, j], limits)
            for ix in range(mc.max_split_tests - 1):
                frac = mc.getFractionIndices(mc.weights, ix + 2)
                for f1, f2 in zip(frac[:-1], frac[1:]):
                    st[j, ix] += (mc.confidence(mc.samples[:, j], limits, start=f1, end=f2) - confids) ** 2
                st[j, ix] = np.sqrt(st[j, ix] / (ix + 2)) / mc.sddev[j]
        out[name + "/split_tests"] = st
    np.savez_compressed(os.path.join(HERE, "f4.npz"), **out)
    print("f4 ok", len(out), "arrays")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES) + ["limits", "f4"]
    if any(nm.startswith("c1") for nm in names) and not os.path.exists(os.path.join(HERE, "c1_inputs.npz")):
        make_c1_inputs()
    for nm in names:
        if nm == "c1_inputs":
            make_c1_inputs()
        elif nm == "limits":
            run_limits()
        elif nm == "f4":
            run_f4()
        else:
            run_case(nm)
