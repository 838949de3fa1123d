"""Seeded synthetic inputs shared by the golden-vector generator (which feeds them to the real
reference) and the tests (which feed them to the oracle and to the CUDA path).

Every case returns a dict: samples (N,P) f64 C-order  | list of per-chain arrays,
weights (N,) f64 | list | None, names, ranges {name: (lo|None, hi|None)}, settings {},
pairs [(jx, jy), ...] 2D densities to evaluate, kwargs_1d / kwargs_2d: list of per-call overrides.
"""

import hashlib

import numpy as np


def _ar1_chol(P, rho):
    R = rho ** np.abs(np.subtract.outer(np.arange(P), np.arange(P)))
    return np.linalg.cholesky(R)


def case_mix3():
    """C1-like: bimodal x0, correlated (x1,x2) at 0.5 (shear branch), Exponential weights."""
    rng = np.random.default_rng(10)
    N = 60000
    comp = rng.random(N) < 0.5
    x0 = np.where(comp, rng.normal(-1.0, 2.0 / 3, N), rng.normal(1.0, 2.0 / 3, N))
    z = rng.normal(size=(N, 2))
    x1 = 3.0 + 0.5 * z[:, 0]
    x2 = -20.0 + 4.0 * (0.5 * z[:, 0] + np.sqrt(0.75) * z[:, 1])
    X = np.ascontiguousarray(np.stack([x0, x1, x2], axis=1))
    w = np.random.default_rng(11).exponential(1.0, N)
    return dict(samples=X, weights=w, names=["a", "b", "c"], ranges={}, settings={},
                pairs=[(0, 1), (0, 2), (1, 2), (2, 1)], kwargs_1d=[{}], kwargs_2d=[{}])


def case_unit5():
    """Unit weights, five nearly independent parameters on very different scales/offsets
    (plain branch with the TNC step; fp32-hazard offsets mean/sigma ~ 150)."""
    rng = np.random.default_rng(21)
    N = 40000
    Z = rng.normal(size=(N, 5)).dot(_ar1_chol(5, 0.15).T)
    sig = np.array([1e-4, 3.0, 50.0, 0.02, 1.0])
    mu = sig * np.array([150.0, -40.0, 0.0, 77.0, 3.0])
    X = np.ascontiguousarray(mu + sig * Z)
    return dict(samples=X, weights=None, names=["p%d" % i for i in range(5)], ranges={}, settings={},
                pairs=[(0, 1), (1, 2), (0, 4), (3, 2)], kwargs_1d=[{}], kwargs_2d=[{}])


def case_bounded():
    """C3-like: hard priors (lower / upper / both / far-away bound that gets dropped)."""
    rng = np.random.default_rng(2025)
    N = 50000
    P = 5
    L = _ar1_chol(P, 0.5)
    lo = np.array([-0.5, -np.inf, -1.5, -np.inf, -30.0])
    hi = np.array([np.inf, 1.0, 1.5, np.inf, np.inf])
    out = np.empty((0, P))
    while out.shape[0] < N:
        Z = rng.normal(size=(N, P)).dot(L.T)
        ok = np.all((Z > lo) & (Z < hi), axis=1)
        out = np.vstack([out, Z[ok]])
    X = np.ascontiguousarray(out[:N] * np.array([1.0, 2.0, 0.5, 1.0, 1.0]) + np.array([0, 0, 0, 5.0, 0]))
    w = np.random.default_rng(2026).exponential(1.0, N)
    ranges = {"q0": (-0.5, None), "q1": (None, 2.0), "q2": (-0.75, 0.75), "q4": (-30.0, None)}
    return dict(samples=X, weights=w, names=["q%d" % i for i in range(P)], ranges=ranges, settings={},
                pairs=[(0, 1), (0, 2), (1, 3), (2, 0), (3, 4), (0, 3)],
                kwargs_1d=[{}, {"boundary_correction_order": 0}, {"boundary_correction_order": 2},
                           {"mult_bias_correction_order": 0}, {"mult_bias_correction_order": 2},
                           {"smooth_scale_1D": 0.3}, {"smooth_scale_1D": 2.0}, {"fine_bins": 512}],
                kwargs_2d=[{}, {"boundary_correction_order": 0, "fine_bins_2D": 128},
                           {"mult_bias_correction_order": 0, "fine_bins_2D": 128},
                           {"smooth_scale_2D": 0.3, "fine_bins_2D": 128},
                           {"smooth_scale_2D": 1.5, "fine_bins_2D": 128},
                           {"mult_bias_correction_order": 2, "fine_bins_2D": 100}])


def case_highcorr():
    """Strong correlations: 0.93 (scaled 512^2 grid, shear), -0.995 (> max_corr_2D, rule of thumb),
    and -0.93."""
    rng = np.random.default_rng(77)
    N = 50000
    z = rng.normal(size=(N, 4))
    x0 = z[:, 0]
    x1 = 0.93 * z[:, 0] + np.sqrt(1 - 0.93**2) * z[:, 1]
    x2 = -0.995 * z[:, 0] + np.sqrt(1 - 0.995**2) * z[:, 2]
    x3 = -0.93 * z[:, 0] + np.sqrt(1 - 0.93**2) * z[:, 3]
    X = np.ascontiguousarray(np.stack([x0, 10 + 3 * x1, x2 * 0.1, x3], axis=1))
    w = 1.0 + np.random.default_rng(78).poisson(2.0, N).astype(np.float64)
    return dict(samples=X, weights=w, names=["h0", "h1", "h2", "h3"], ranges={}, settings={},
                pairs=[(0, 1), (0, 2), (0, 3), (1, 3)], kwargs_1d=[{}], kwargs_2d=[{}])


def case_chains():
    """C4-like: 4 chains, integer weights, mean-shifted chains -> Gelman-Rubin + cov."""
    rng = np.random.default_rng(777)
    P = 6
    L = _ar1_chol(P, 0.7)
    chains, weights, loglikes = [], [], []
    for c in range(4):
        n = 9000 + 500 * c
        Z = rng.normal(size=(n, P)).dot(L.T) + 0.05 * rng.normal(size=P)
        chains.append(np.ascontiguousarray(Z * np.array([1, 10, 0.1, 1, 1, 3.0]) + np.array([0, 100, 0, -5, 0, 0])))
        weights.append(1.0 + rng.poisson(2.0, n).astype(np.float64))
        loglikes.append(0.5 * np.sum(Z**2, axis=1))  # the reference's getSeparateChains needs loglikes
    return dict(samples=chains, weights=weights, loglikes=loglikes, names=["g%d" % i for i in range(P)], ranges={}, settings={},
                pairs=[(0, 1), (2, 5)], kwargs_1d=[{}], kwargs_2d=[{}])


def case_mcmc():
    """Metropolis-like chain: AR(1) proposals with rejections folded into integer multiplicities, three
    parameters with short / long / no autocorrelation -> the default sampler="mcmc" N_eff path."""
    rng = np.random.default_rng(4242)
    n = 30000
    x = np.empty((n, 3))
    e = rng.normal(size=(n, 3))
    x[0] = e[0]
    phi = np.array([0.6, 0.97, 0.0])
    for i in range(1, n):
        x[i] = phi * x[i - 1] + np.sqrt(1 - phi**2) * e[i]
    w = 1.0 + rng.geometric(0.4, n).astype(np.float64)
    X = np.ascontiguousarray(x * np.array([1.0, 5.0, 0.2]) + np.array([0.0, 50.0, 1.0]))
    return dict(samples=X, weights=w, names=["m0", "m1", "m2"], ranges={}, settings={}, sampler="mcmc",
                pairs=[(0, 1), (0, 2)], kwargs_1d=[{}], kwargs_2d=[{}])


def case_periodic():
    """testPeriodic-like (getdist_test.py:181-225): a wrapped angle (periodic), a bounded radius, a free parameter
    and a second angle -> 1D periodic, 2D periodic in x, in y and in both."""
    rng = np.random.default_rng(42)
    n = 30000
    angle = rng.normal(0, 1, n) % (2 * np.pi)
    radius = np.abs(rng.normal(2, 0.5, n))
    free = rng.normal(size=n) + 0.3 * np.cos(angle)
    phase = (rng.vonmises(1.0, 2.0, n) + 0.2 * free) % (2 * np.pi)
    X = np.ascontiguousarray(np.column_stack([angle, radius, free, phase]))
    w = np.random.default_rng(43).exponential(1.0, n)
    ranges = {"angle": [0, 2 * np.pi, "periodic"], "radius": [0, 5], "phase": [0, 2 * np.pi, True]}
    return dict(samples=X, weights=w, names=["angle", "radius", "free", "phase"], ranges=ranges, settings={},
                pairs=[(0, 1), (1, 0), (0, 2), (0, 3), (2, 3)],
                kwargs_1d=[{}, {"mult_bias_correction_order": 0}, {"fine_bins": 64}],
                kwargs_2d=[{}, {"fine_bins_2D": 64}, {"mult_bias_correction_order": 0, "fine_bins_2D": 64}])


def likes_mask(minx, miny, stepx, stepy, mask):
    """mask_function of case_likes (signature of mcsamples.py:1909-1916): excludes the half plane x + 0.5 y > 4, a
    2.2 sigma tail.  (A mask cutting through the bulk of the SAMPLES makes the reference's linear boundary
    correction divide noise by noise inside the masked region: its own result then moves by 1e-5 .. 1e-2 with the
    FFT size, measured; priors that the samples obey, or tail cuts like this one, are stable at 1e-15.)"""
    ny, nx = mask.shape
    x = minx + stepx * np.arange(nx)
    y = miny + stepy * np.arange(ny)
    mask[(x[None, :] + 0.5 * y[:, None]) > 4.0] = 0


def case_likes():
    """meanlikes / mask_function: correlated Gaussian with its true -log(likelihood), one parameter
    cut by a hard prior, Exponential weights; 1D mean likelihoods, 2D mean likelihoods (with and without the bias
    correction) and a prior mask on pair (0, 1)."""
    rng = np.random.default_rng(909)
    N, P = 40000, 4
    L = _ar1_chol(P, 0.6)
    out = np.empty((0, P))
    while out.shape[0] < N:
        Z = rng.normal(size=(N, P)).dot(L.T)
        out = np.vstack([out, Z[Z[:, 2] > -0.8]])
    Z = out[:N]
    Rinv = np.linalg.inv(L.dot(L.T))
    loglikes = 0.5 * np.einsum("ni,ij,nj->n", Z, Rinv, Z) + 3.7
    X = np.ascontiguousarray(Z * np.array([1.0, 2.0, 0.5, 30.0]) + np.array([0.0, 0.0, 0.0, 400.0]))
    w = np.random.default_rng(910).exponential(1.0, N)
    return dict(samples=X, weights=w, loglikes=loglikes, names=["l0", "l1", "l2", "l3"], ranges={"l2": (-0.4, None)},
                settings={}, pairs=[(0, 1), (1, 2), (3, 2)], kwargs_1d=[{}], kwargs_2d=[{}],
                meanlikes=True, likes_kwargs_1d=[{}, {"mult_bias_correction_order": 0}, {"smooth_scale_1D": 0.4}],
                likes_kwargs_2d=[{}, {"mult_bias_correction_order": 0}, {"smooth_scale_2D": 0.5, "fine_bins_2D": 128}],
                mask_function=likes_mask, mask_pairs=[(0, 1)],
                mask_kwargs_2d=[{}, {"mult_bias_correction_order": 0, "fine_bins_2D": 128}])


def _c1_inputs():
    """C1 (BASELINE.json configs[0]): samples drawn by the REFERENCE's own mixture classes
    (getdist.gaussian_mixtures, make_golden.py:make_c1_inputs) and stored as a fixture -- the reference is not importable
    where the GPU tests run."""
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_inputs.npz"))


def case_c1rand():
    """RandomTestMixtureND(ndim=3, ncomponent=2, seed=10).MCSamples(100000, random_state=10), unit weights"""
    z = _c1_inputs()
    return dict(samples=np.ascontiguousarray(z["rand"]), weights=None, names=["x", "y", "z"], ranges={}, settings={},
                pairs=[(0, 1), (0, 2), (1, 2)], kwargs_1d=[{}], kwargs_2d=[{}])


def case_c1rand_w():
    """the same samples with Exponential(1) weights (seed 11)"""
    z = _c1_inputs()
    return dict(samples=np.ascontiguousarray(z["rand"]), weights=np.ascontiguousarray(z["w"]), names=["x", "y", "z"], ranges={},
                settings={}, pairs=[(0, 1), (0, 2), (1, 2)], kwargs_1d=[{}], kwargs_2d=[{}])


def case_c1wj():
    """bimodal WJ1 mixture Mixture2D([[-1,0],[1,0]], [(2/3,2/3,0)]*2) (test_distributions.py:196-198) with a third
    Gaussian dimension, 100000 samples, unit weights"""
    z = _c1_inputs()
    return dict(samples=np.ascontiguousarray(z["wj"]), weights=None, names=["x", "y", "z"], ranges={}, settings={},
                pairs=[(0, 1), (0, 2), (1, 2)], kwargs_1d=[{}], kwargs_2d=[{}])


CASES = {
    "c1rand": case_c1rand,
    "c1rand_w": case_c1rand_w,
    "c1wj": case_c1wj,
    "mix3": case_mix3,
    "unit5": case_unit5,
    "bounded": case_bounded,
    "highcorr": case_highcorr,
    "chains": case_chains,
    "mcmc": case_mcmc,
    "periodic": case_periodic,
    "likes": case_likes,
}


def input_digest(case):
    h = hashlib.sha256()
    s = case["samples"]
    for a in (s if isinstance(s, list) else [s]):
        h.update(np.ascontiguousarray(a).tobytes())
    w = case["weights"]
    if w is not None:
        for a in (w if isinstance(w, list) else [w]):
            h.update(np.ascontiguousarray(a).tobytes())
    if case.get("meanlikes"):  # cases whose goldens depend on the log-likelihoods
        h.update(np.ascontiguousarray(case["loglikes"]).tobytes())
    return h.hexdigest()


def grid_stride(size):
    """Golden 2D grids larger than 256^2 are stored on a strided subsample to keep fixtures small."""
    return 1 if size <= 256 else max(1, size // 128)


def kw_tag(kw):
    return "default" if not kw else ",".join("%s=%s" % (k, kw[k]) for k in sorted(kw))
