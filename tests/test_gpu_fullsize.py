"""GPU tests at (or near) BASELINE.json's full sizes, through size-independent properties plus oracle parity on a
bounded column sample: C3 (hard-bounded priors, N=5e6, P=32) and C4 (4 chains, P=128: moments + Gelman-Rubin)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ar1(P, rho):
    return np.linalg.cholesky(rho ** np.abs(np.subtract.outer(np.arange(P), np.arange(P))))


def test_c3_bounded_full_size():
    from getdist_b200 import MCSamples
    from oracle.getdist_oracle import OracleSamples

    rng = np.random.default_rng(2025)
    N, P = 5_000_000, 32
    L = _ar1(P, 0.5)
    lo = np.full(P, -np.inf)
    hi = np.full(P, np.inf)
    lo[0:8] = -0.5
    hi[8:16] = 1.0
    lo[16:20], hi[16:20] = -1.5, 1.5
    X = np.empty((0, P))
    while X.shape[0] < N:
        Z = rng.normal(size=(N // 2, P)).dot(L.T)
        X = np.vstack([X, Z[np.all((Z > lo) & (Z < hi), axis=1)]])
    X = np.ascontiguousarray(X[:N])
    w = rng.exponential(1.0, N)
    names = ["q%d" % i for i in range(P)]
    ranges = {names[i]: (None if np.isinf(lo[i]) else lo[i], None if np.isinf(hi[i]) else hi[i]) for i in range(P)
              if not (np.isinf(lo[i]) and np.isinf(hi[i]))}
    mc = MCSamples(samples=X, weights=w, names=names, ranges=ranges, sampler="uncorrelated")
    d1, d2 = mc.prefetch_triangle()
    assert len(d1) == P and len(d2) == P * (P - 1) // 2
    # properties: max-normalised, non-negative, finite; bounded parameters keep their hard range
    for d in d1:
        assert np.isfinite(d.P).all() and d.P.min() >= 0 and abs(d.P.max() - 1) < 1e-15
    for d in d2:
        assert np.isfinite(d.P).all() and d.P.min() >= -1e-12 and abs(d.P.max() - 1) < 1e-15
    assert mc.paramNames.names[0].range_min == -0.5 and mc.paramNames.names[8].range_max == 1.0
    # 2D histogram mass is conserved for every pair of a tile-crossing subset
    pairs = [(0, 1), (7, 8), (3, 31), (16, 17), (15, 24)]
    specs = [mc._spec_2d(a, b, {}) for (a, b) in pairs]
    buf, offs = mc._ctx.hist2d_batch(specs)
    for sp, off in zip(specs, offs):
        np.testing.assert_allclose(buf[off: off + sp.fine_bins ** 2].sum(), w.sum(), rtol=1e-12)
    # oracle parity on a column sample (bounded x bounded, bounded x free, free x free) at full N
    cols = [0, 9, 17, 30]
    o = OracleSamples(np.ascontiguousarray(X[:, cols]), w, names=[names[c] for c in cols],
                      ranges={names[c]: ranges[names[c]] for c in cols if names[c] in ranges})
    for k, c in enumerate(cols):
        assert np.max(np.abs(mc.get1DDensity(names[c]).P - o.density_1d(k).P)) < 1e-6
    for (a, b) in [(0, 1), (1, 2), (2, 3)]:
        d = mc.get2DDensity(names[cols[a]], names[cols[b]])
        tol = 1e-5 if d._gdk["status"] & (64 | 128) else 1e-6
        assert np.max(np.abs(d.P - o.density_2d(a, b).P)) < tol, (a, b)


def test_c4_chain_statistics():
    from getdist_b200 import MCSamples
    from oracle.getdist_oracle import gelman_rubin, weighted_cov, weighted_means

    rng = np.random.default_rng(77)
    P, n = 128, 600_000
    L = _ar1(P, 0.7)
    sig = 10.0 ** rng.uniform(-2, 2, P)
    chains, ws = [], []
    for c in range(4):
        Z = rng.normal(size=(n, P)).dot(L.T) + 0.01 * rng.normal(size=P)
        chains.append(np.ascontiguousarray(Z * sig + 100 * sig))
        ws.append(1.0 + rng.poisson(2.0, n).astype(np.float64))
    mc = MCSamples(samples=chains, weights=ws, sampler="uncorrelated")
    X, w = np.vstack(chains), np.hstack(ws)
    m = weighted_means(X, w)
    np.testing.assert_allclose(mc.getMeans(), m, rtol=1e-12)
    cov = weighted_cov(X, w, m, blocked=True)
    scale = np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
    assert np.max(np.abs(mc.getCov() - cov) / scale) < 1e-11
    offs = np.cumsum([0] + [c.shape[0] for c in chains])
    # the whitening in getGelmanRubinEigenvalues amplifies rounding by the condition number of the mean covariance
    np.testing.assert_allclose(mc.getGelmanRubin(64), gelman_rubin(X[:, :64], w, offs, 64), rtol=1e-6)
    # properties: symmetric PSD covariance, chain order does not matter
    C = mc.getCov()
    assert np.array_equal(C, C.T) and np.linalg.eigvalsh(C).min() > 0
    mc2 = MCSamples(samples=chains[::-1], weights=ws[::-1], sampler="uncorrelated")
    np.testing.assert_allclose(mc2.getGelmanRubin(64), mc.getGelmanRubin(64), rtol=1e-8)
    np.testing.assert_allclose(mc2.getCov(), C, rtol=1e-10, atol=1e-12 * np.abs(C).max())


def test_c4_full_size_chains():
    """BASELINE.json configs[3] at its own size: 4 chains x 2.5e6 rows x 128 parameters.  Moments and
    Gelman-Rubin from the fused one-sweep statistics, against the oracle on a column sample, plus properties that hold
    whatever the size: the statistics do not depend on how the rows were chunked (re-run of the sweep == the sweep that
    rode behind the upload, bit for bit) and chain order does not matter."""
    from getdist_b200 import MCSamples
    from oracle.getdist_oracle import gelman_rubin, weighted_cov, weighted_means

    rng = np.random.default_rng(77)
    P, n = 128, 2_500_000
    L = _ar1(P, 0.7)
    sig = 10.0 ** rng.uniform(-2, 2, P)
    chains, ws = [], []
    for c in range(4):
        Z = rng.standard_normal((n, P)).dot(L.T) + 0.01 * rng.normal(size=P)
        chains.append(np.ascontiguousarray(Z * sig + 100 * sig))
        ws.append(1.0 + rng.poisson(2.0, n).astype(np.float64))
    mc = MCSamples(samples=chains, weights=ws, sampler="uncorrelated")
    cols = list(range(0, P, 8))  # 16 columns
    Xs = np.vstack([c[:, cols] for c in chains])
    w = np.hstack(ws)
    m = weighted_means(Xs, w)
    np.testing.assert_allclose(mc.getMeans()[cols], m, rtol=1e-12)
    cov = weighted_cov(Xs, w, m, blocked=True)
    scale = np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
    assert np.max(np.abs(mc.getCov()[np.ix_(cols, cols)] - cov) / scale) < 1e-11
    offs = np.cumsum([0] + [c.shape[0] for c in chains])
    np.testing.assert_allclose(mc.getGelmanRubin(8), gelman_rubin(np.vstack([c[:, :8] for c in chains]), w, offs, 8), rtol=1e-6)
    C = mc.getCov().copy()
    assert np.array_equal(C, C.T) and np.linalg.eigvalsh(C).min() > 0
    # exact column extrema (the quantile histograms are laid out on them)
    assert np.array_equal(mc._xmin[cols], Xs.min(0)) and np.array_equal(mc._xmax[cols], Xs.max(0))
    # the sweep run again on the resident store gives the same bits as the one pipelined behind the upload chunks
    mc._ctx.moments_recompute()
    m2 = mc._ctx.moments()
    assert np.array_equal(m2["cov"], C) and np.array_equal(m2["means"], mc.getMeans())


def test_c2_rho095_variable_grids():
    """SURVEY.md s8d, C2 variant rho = 0.95: neighbouring parameters are correlated strongly enough for the scaled
    fine grids (mcsamples.py:1811-1818: 384 .. 960 bins) and the sheared bandwidth branch; these grids leave the
    256^2 fast path.  Oracle parity on pairs of every grid size that occurs, properties on the rest."""
    from getdist_b200 import MCSamples
    from oracle.getdist_oracle import OracleSamples

    rng = np.random.default_rng(1234)
    N, P = 1_000_000, 12
    X = rng.standard_normal((N, P)).dot(_ar1(P, 0.95).T) * 10.0 ** rng.uniform(-3, 2, P) + rng.uniform(-50, 50, P)
    w = rng.exponential(1.0, N)
    names = ["p%d" % i for i in range(P)]
    settings = {"fine_bins": 2048, "fine_bins_2D": 256}
    mc = MCSamples(samples=X, weights=w, names=names, sampler="uncorrelated", settings=settings)
    d1, d2 = mc.prefetch_triangle()
    idx, pairs = mc.triangle_pairs()
    sizes = sorted({d.P.shape[0] for d in d2})
    assert sizes[0] == 256 and sizes[-1] > 256, sizes
    for d in d2:
        assert d.P.shape[0] == d.P.shape[1] and np.isfinite(d.P).all() and abs(d.P.max() - 1) < 1e-15
    by_size = {}
    for k, d in enumerate(d2):
        by_size.setdefault(d.P.shape[0], []).append(k)
    for G, ks in by_size.items():
        k = ks[len(ks) // 2]
        a, b = pairs[k]
        o = OracleSamples(np.ascontiguousarray(X[:, [a, b]]), w, names=[names[a], names[b]], sampler="uncorrelated", settings=settings)
        ref = o.density_2d(0, 1)
        assert ref.P.shape == d2[k].P.shape, (G, ref.P.shape)
        tol = 1e-5 if d2[k]._gdk["status"] & (64 | 128) else 1e-6
        assert np.max(np.abs(d2[k].P - ref.P)) < tol, (G, a, b, np.max(np.abs(d2[k].P - ref.P)))
