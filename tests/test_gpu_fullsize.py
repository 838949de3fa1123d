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
