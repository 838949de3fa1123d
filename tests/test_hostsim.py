"""CPU tests of the grid-stage DEVICE arithmetic: the .cuh headers the CUDA kernels are built from
are compiled for the host (tests/hostsim, one-thread cooperative group) and compared with numpy/scipy
and with the oracle.  Checks the math, not the CUDA execution (that is what the -m gpu tests do)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy import fftpack
from scipy.optimize import brentq, fsolve

from cases import kw_tag
from helpers import load_case, make_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="session")
def hs():
    src = os.path.join(HERE, "hostsim", "hostsim.cpp")
    out = os.path.join(HERE, "hostsim", "libhostsim.so")
    deps = [src] + [os.path.join(ROOT, "getdist_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "getdist_b200", "csrc"))
                    if f.endswith((".cuh", ".h"))] + [os.path.join(ROOT, "include", "gdk.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", out, src])
    return C.CDLL(out)


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 100, 384, 7])
def test_fft_dct(hs, n):
    rng = np.random.default_rng(n)
    nl = 3
    x = rng.normal(size=(nl, n)) + 1j * rng.normal(size=(nl, n))
    re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    ore, oim = np.empty_like(re), np.empty_like(im)
    hs.hs_fft(dptr(re), dptr(im), n, nl, dptr(ore), dptr(oim))
    ref = np.fft.fft(x, axis=1)
    assert np.max(np.abs(ore + 1j * oim - ref)) < 1e-13 * n * np.max(np.abs(ref))
    r = np.ascontiguousarray(rng.random((nl, n)))
    out = np.empty_like(r)
    hs.hs_dct2(dptr(r), n, nl, dptr(out))
    ref = fftpack.dct(r, axis=1)
    assert np.max(np.abs(out - ref)) < 2e-15 * n * np.max(np.abs(ref))


CUBICS = [(1.0, 0.5, 0.3), (5.0, 0.01, 0.02), (-2.0, -1.0, 0.7), (100.0, 3.0, 0.0123), (0.3, 2.0, 0.11)]


@pytest.mark.parametrize("a,b,r", CUBICS)
def test_brentq_port_matches_scipy(hs, a, b, r):
    pts = []

    def f(x):
        pts.append(x)
        d = x - r
        return a * d * d * d + b * d

    for (xa, xb, xtol) in [(0.0, 1.0, 1e-6), (r - 0.4, r + 0.9, 1e-3), (0.0, 0.1 + r, 1e-12)]:
        pts.clear()
        ref = brentq(f, xa, xb, xtol=xtol)
        xs = np.zeros(512)
        xo = C.c_double()
        nf = C.c_int()
        st = hs.hs_brentq(C.c_double(a), C.c_double(b), C.c_double(r), C.c_double(xa), C.c_double(xb),
                          C.c_double(xtol), C.byref(xo), dptr(xs), C.byref(nf))
        assert st == 0
        assert xo.value == ref
        assert nf.value == len(pts)
        assert np.array_equal(xs[: nf.value], np.array(pts))


@pytest.mark.parametrize("a,b,r", CUBICS)
def test_hybrd_port_matches_fsolve(hs, a, b, r):
    pts = []

    def f(x):
        pts.append(float(x[0]))
        d = x[0] - r
        return [a * d * d * d + b * d]

    for x0 in [0.53 * r + 0.01, 1.7 * r + 0.05, 0.2]:
        pts.clear()
        ref = fsolve(f, x0, xtol=x0 / 20, factor=1)[0]
        xs = np.zeros(512)
        xo = C.c_double()
        nf = C.c_int()
        hs.hs_hybrd1(C.c_double(a), C.c_double(b), C.c_double(r), C.c_double(x0), C.c_double(x0 / 20),
                     C.c_double(1.0), C.byref(xo), dptr(xs), C.byref(nf))
        # scipy makes extra bookkeeping evaluations at x0 (shape check etc.): compare the iterates
        # away from x0 (x0 itself and the forward-difference point x0*(1+1.5e-8) are dropped)
        ours = xs[: nf.value]
        theirs = np.array(pts)
        ours = ours[np.abs(ours - x0) > 1e-6 * abs(x0)]
        theirs = theirs[np.abs(theirs - x0) > 1e-6 * abs(x0)]
        assert len(ours) == len(theirs), (ours, theirs)
        np.testing.assert_allclose(ours, theirs, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(xo.value, ref, rtol=1e-12)


class Spec1D(C.Structure):
    _fields_ = [("param", C.c_int32), ("fine_bins", C.c_int32), ("binmin", C.c_double), ("binmax", C.c_double),
                ("range_min", C.c_double), ("range_max", C.c_double), ("param_min", C.c_double),
                ("param_max", C.c_double), ("sigma_range", C.c_double), ("err", C.c_double), ("neff", C.c_double),
                ("smooth_scale_1D", C.c_double), ("width", C.c_double), ("boundary_correction_order", C.c_int32),
                ("mult_bias_correction_order", C.c_int32), ("has_limits_bot", C.c_int32), ("has_limits_top", C.c_int32),
                ("periodic", C.c_int32), ("pad", C.c_int32)]


class Res1D(C.Structure):
    _fields_ = [("kde_h", C.c_double), ("h_raw", C.c_double), ("smooth_1D", C.c_double), ("winw", C.c_int32),
                ("status", C.c_uint32), ("n_feval", C.c_int32), ("pad", C.c_int32)]


@pytest.mark.parametrize("name", ["mix3", "unit5", "bounded", "highcorr", "chains", "mcmc", "periodic"])
def test_kde1d_core_vs_oracle_and_golden(hs, name):
    from oracle.getdist_oracle import bin_geometry, bin_indices

    case, g = load_case(name)
    o = make_oracle(case)
    for kw in case["kwargs_1d"]:
        tag = kw_tag(kw)
        for j in range(o.n):
            d = o.density_1d(j, **kw)
            par = o.pars[j]
            s = dict(o.settings)
            s.update(kw)
            F = s["fine_bins"]
            binmin, binmax, fw = bin_geometry(par, F)
            bins = np.bincount(bin_indices(o.samples[:, j], binmin, fw), weights=o.weights, minlength=F)
            sp = Spec1D(j, F, binmin, binmax, par.range_min, par.range_max, par.param_min, par.param_max,
                        par.sigma_range, par.err, o._neff(par), s["smooth_scale_1D"],
                        (par.range_max - par.range_min) / (s["num_bins"] - 1), s["boundary_correction_order"],
                        s["mult_bias_correction_order"], int(par.has_limits_bot), int(par.has_limits_top), int(par.periodic), 0)
            P = np.empty(F)
            res = Res1D()
            hs.hs_kde1d(C.byref(sp), dptr(bins), dptr(P), C.byref(res))
            assert res.winw == d.winw, (name, tag, j)
            if s["smooth_scale_1D"] <= 0:
                np.testing.assert_allclose(res.kde_h, d.h, rtol=2e-6)
            err = np.max(np.abs(P - d.P))
            gerr = np.max(np.abs(P - g["d1/%s/%d/P" % (tag, j)]))
            assert err < 1e-7 and gerr < 1e-7, (name, tag, j, err, gerr)


def _pair_inputs(o, jx, jy, G=256):
    from oracle.getdist_oracle import bin_geometry, bin_indices

    px, py = o.init_param_ranges(jx), o.init_param_ranges(jy)
    xb, yb = bin_geometry(px, G), bin_geometry(py, G)
    H = o._hist2d(bin_indices(o.samples[:, jx], xb[0], xb[2]), bin_indices(o.samples[:, jy], yb[0], yb[2]), G, G)
    return px, py, xb, yb, H


@pytest.mark.parametrize("name,jx,jy,G", [("mix3", 0, 1, 256), ("unit5", 1, 2, 256), ("unit5", 0, 4, 256),
                                         ("bounded", 0, 1, 256), ("bounded", 0, 2, 100), ("unit5", 3, 2, 128),
                                         ("c1rand_w", 1, 2, 256)])
def test_bw2d_core_vs_oracle(hs, name, jx, jy, G):
    """2D transforms + KernelOptimizer2D restatement (plain branch) against the oracle's optimiser."""
    from oracle.getdist_oracle import BandwidthOptimizer2D

    case, g = load_case(name)
    o = make_oracle(case)
    px, py, xb, yb, H = _pair_inputs(o, jx, jy, G)
    neff = min(o._neff(px), o._neff(py))
    corr = o.get_correlation_matrix()[jy][jx]
    do_corr = not (px.has_limits or py.has_limits)
    ft = (min(py.sigma_range / (yb[1] - yb[0]), px.sigma_range / (xb[1] - xb[0])) / neff ** (1.0 / 6)) ** 2
    ref = BandwidthOptimizer2D(H, neff, corr, do_correlation=do_corr, fallback_t=ft)
    a2 = np.empty((G, G))
    aF = np.empty((G, G))
    hs.hs_xform2d(dptr(np.ascontiguousarray(H)), G, dptr(a2), dptr(aF))
    assert np.max(np.abs(a2[1:, 1:] - ref.a2)) < 1e-13 * np.max(ref.a2)
    if do_corr:
        assert np.max(np.abs(aF - ref.aFFT.real)) < 1e-13 * np.max(ref.aFFT.real)
    out = np.zeros(4)
    iout = np.zeros(3, dtype=np.int32)
    hs.hs_bw2d(dptr(a2), dptr(aF), G, C.c_double(neff), C.c_double(corr), int(do_corr), 1, C.c_double(ft), dptr(out),
               iout.ctypes.data_as(C.POINTER(C.c_int)))
    assert iout[2] == 0
    assert iout[1] == ref.n_brent_evals  # same Brent path
    np.testing.assert_allclose(out[3], ref.t_star, rtol=1e-9)
    hx, hy, c = ref.get_h()
    np.testing.assert_allclose(ref.h_closed, ref.h_closed)
    # closed-form part is tight; where the reference's TNC result is accepted, h scatters by ~1e-4
    tnc = (c != 0) or (abs(hx - ref.h_closed[0]) > 0)
    rtol = 5e-4 if tnc else 1e-9
    np.testing.assert_allclose(out[:2], [hx, hy], rtol=rtol)
    np.testing.assert_allclose(out[2], c, rtol=1e-12, atol=1e-15)
    if tnc and c != 0:
        # where the reference's TNC decides the widths it stops (on finite-difference gradients) short of the minimum of
        # the AMISE; the safeguarded Newton iteration of the device code ends at or below the reference's value
        assert ref.amise(np.array([out[0], out[1]]), c) <= ref.amise(np.array([hx, hy]), c) * (1 + 1e-12)


def test_contour_levels_core(hs):
    """device contour-level bisection vs the sort-based getContourLevels of the reference (golden 2D grids)"""
    from getdist_b200.densities import getContourLevels

    conts = np.array([0.68, 0.95, 0.99])
    for name in ["mix3", "bounded", "chains"]:
        case, g = load_case(name)
        for (jx, jy) in case["pairs"]:
            P = np.ascontiguousarray(g["d2/default/%d_%d/P" % (jx, jy)])
            if P.shape[0] != 256:
                continue
            ref = getContourLevels(P, conts)
            lv = np.zeros(4)
            out = hs.hs_contours(dptr(P), P.shape[0], dptr(conts), 3, dptr(lv))
            assert out == 0
            np.testing.assert_allclose(lv[:3], ref, rtol=1e-9, atol=1e-14)


def test_kde1d_core_meanlikes(hs):
    """mean likelihoods of the 1D grid stage (kde1d_core, mcsamples.py:1556-1561, 1597-1598, 1672-1684) against the
    oracle and the reference golden (case 'likes': unbounded, bounded, and fixed-width kernels)"""
    from oracle.getdist_oracle import bin_geometry, bin_indices

    case, g = load_case("likes")
    o = make_oracle(case)
    lw = o.weights * np.exp(o.mean_loglike - o.loglikes)
    for kw in case["likes_kwargs_1d"]:
        tag = kw_tag(kw)
        for j in range(o.n):
            d = o.density_1d(j, meanlikes=True, **kw)
            par = o.pars[j]
            s = dict(o.settings)
            s.update(kw)
            F = s["fine_bins"]
            binmin, binmax, fw = bin_geometry(par, F)
            ix = bin_indices(o.samples[:, j], binmin, fw)
            bins = np.bincount(ix, weights=o.weights, minlength=F)
            lbins = np.bincount(ix, weights=lw, minlength=F)
            sp = Spec1D(j, F, binmin, binmax, par.range_min, par.range_max, par.param_min, par.param_max,
                        par.sigma_range, par.err, o._neff(par), s["smooth_scale_1D"],
                        (par.range_max - par.range_min) / (s["num_bins"] - 1), s["boundary_correction_order"],
                        s["mult_bias_correction_order"], int(par.has_limits_bot), int(par.has_limits_top), int(par.periodic), 0)
            P, L = np.empty(F), np.empty(F)
            res = Res1D()
            hs.hs_kde1d_likes(C.byref(sp), dptr(bins), dptr(lbins), dptr(P), dptr(L), C.byref(res))
            assert np.max(np.abs(P - d.P)) < 1e-7
            assert np.max(np.abs(L - d.likes)) < 1e-6, (tag, j, np.max(np.abs(L - d.likes)))
            assert np.max(np.abs(L - g["l1/%s/%d/likes" % (tag, j)])) < 1e-6, (tag, j)


# ---------------------------------------------------------------------------------------------------------------
# barrier structure of the grid-stage device code: the same headers on REAL host threads under ThreadSanitizer
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="session")
def race_check(tmp_path_factory):
    src = os.path.join(HERE, "hostsim", "race_check.cpp")
    exe = str(tmp_path_factory.mktemp("race") / "race_check")
    probe = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-o", exe, src],
                           capture_output=True, text=True, cwd=os.path.join(HERE, "hostsim"))
    if probe.returncode != 0:
        if "tsan" in probe.stderr.lower() or "sanitize" in probe.stderr.lower():
            pytest.skip("no ThreadSanitizer runtime for this g++")
        raise RuntimeError(probe.stderr[-2000:])
    return exe


def _run_tsan(exe, *args):
    env = dict(os.environ, TSAN_OPTIONS="exitcode=66 halt_on_error=0")
    return subprocess.run([exe, *args], capture_output=True, text=True, env=env, timeout=600)


@pytest.mark.parametrize("threads", [32, 64, 128])
def test_device_code_barriers_under_thread_sanitizer(race_check, threads):
    """fft / dct lines, the whole 1D grid stage (k_kde1d's body: every boundary / bias / periodic / likes variant), the
    2D bandwidth optimiser (k_bw2d's body) and the contour selection (k_contours2d's body), instantiated with a group
    of real threads whose co.sync() is a pthread barrier: no data race (a missing barrier between a shared-scratch
    write and another thread's read would be one), and the same numbers as the one-thread instantiation"""
    r = _run_tsan(race_check, str(threads))
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0, r.stdout[-3000:]
    assert "all cases agree" in r.stdout and "MISMATCH" not in r.stdout


def test_thread_sanitizer_sees_a_missing_barrier(race_check):
    """the checker checks itself: with the barriers turned into no-ops the FFT stages race and ThreadSanitizer says so"""
    r = _run_tsan(race_check, "64", "--break-barriers")
    assert r.returncode == 66 and "data race" in r.stderr and "fft.cuh" in r.stderr
