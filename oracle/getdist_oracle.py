"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A plain numpy/scipy restatement of the GetDist (cmbant/getdist 1.7.7) hot path that
``getdist_b200`` re-implements in CUDA: weighted moments / covariance / Gelman-Rubin,
exact weighted order statistics, the 1D and 2D FFT-KDE pipelines (fine-grid weighted
histogram, Botev improved-Sheather-Jones bandwidth in 1D and 2D, Gaussian-kernel
convolution, linear boundary correction, multiplicative bias correction, max-normalisation).

Nothing in the product path (``getdist_b200/``) may import this module.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker / the timed CPU baseline.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` in the build container and stores its outputs (means, cov, Gelman-Rubin,
per-parameter ranges, bandwidths, 1D and 2D density grids) under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.

Third-party arithmetic the reference leans on is called here through the same libraries
(numpy ``bincount/argsort/cumsum/searchsorted/convolve``, ``scipy.fftpack.dct``,
``scipy.optimize.fsolve/brentq/minimize(TNC)``), as the reference's own call sites do
(kde_bandwidth.py:116,123,127,162,276-299; mcsamples.py:1554,1728; chains.py:806-811,836).

Every function cites the reference file:line it restates (paths relative to
``/root/reference/getdist/``).  The cost structure deliberately mirrors the reference (e.g. the
weighted quantiles are recomputed by a full argsort for every 1D density and twice for every
2D density, mcsamples.py:1541,1786-1787) because ``bench.py --impl reference`` times this code
as the CPU arm.
"""

from __future__ import annotations

import logging
import math
import warnings
from dataclasses import dataclass, field

import numpy as np
from scipy import fftpack
from scipy.optimize import brentq, fsolve, minimize
from scipy.signal import fftconvolve

log = logging.getLogger("getdist_oracle")


class OracleError(Exception):
    pass


class OracleSettingError(OracleError):
    pass


class OracleBandwidthError(OracleError):
    pass


class OracleDensityError(OracleError):
    pass


# --------------------------------------------------------------------------------------
# settings (analysis_defaults.ini values, which override the class attributes -- SURVEY s5)
# --------------------------------------------------------------------------------------
DEFAULT_SETTINGS = dict(
    fine_bins=1024,
    fine_bins_2D=256,
    smooth_scale_1D=-1.0,
    smooth_scale_2D=-1.0,
    boundary_correction_order=1,
    mult_bias_correction_order=1,
    max_corr_2D=0.99,
    range_confidence=0.001,
    num_bins=100,
    num_bins_2D=40,
    use_effective_samples_2D=False,
    range_ND_contour=-1,
)


@dataclass
class ParamState:
    """Per-parameter state the reference keeps on ``ParamInfo`` (mcsamples.py:1427-1484, 442-470)."""

    name: str
    limmin: float | None = None
    limmax: float | None = None
    periodic: bool = False
    has_limits_bot: bool = False
    has_limits_top: bool = False
    has_limits: bool = False
    err: float = 0.0
    mean: float = 0.0
    param_min: float = 0.0
    param_max: float = 0.0
    range_min: float = 0.0
    range_max: float = 0.0
    sigma_range: float = 0.0
    N_eff_kde: float | None = None
    kde_h: float | None = None


@dataclass
class Grid1D:
    x: np.ndarray
    P: np.ndarray
    view_ranges: tuple
    h: float | None = None  # bandwidth fraction actually used (kde_h after fallback), for diagnostics
    winw: int = 0
    smooth_bins: float = 0.0
    likes: np.ndarray | None = None  # mean likelihoods (meanlikes=True), mcsamples.py:1672-1684


@dataclass
class Grid2D:
    x: np.ndarray
    y: np.ndarray
    P: np.ndarray
    view_ranges: tuple
    rx: float = 0.0
    ry: float = 0.0
    corr: float = 0.0
    winw: int = 0
    fine_bins: int = 0
    extra: dict = field(default_factory=dict)
    likes: np.ndarray | None = None  # mean likelihoods (meanlikes=True, get_density=False), mcsamples.py:2004-2006
    mask: np.ndarray | None = None   # bool_mask of a mask_function, mcsamples.py:1917, 1986


# --------------------------------------------------------------------------------------
# weighted statistics
# --------------------------------------------------------------------------------------
def weighted_means(X, w):
    """chains.py:373-384: ``means = weights.dot(samples) / norm``."""
    return w.dot(X) / np.sum(w)


def weighted_vars(X, w, means):
    """chains.py:400-412: per-parameter ``weights.dot((x - mean)**2) / norm``."""
    norm = np.sum(w)
    out = np.empty(X.shape[1])
    for i in range(X.shape[1]):
        out[i] = w.dot((X[:, i] - means[i]) ** 2) / norm
    return out


def weighted_cov(X, w, means=None, blocked=False):
    """chains.py:709-733 (``cov``): two-pass centred covariance, population normalisation.

    ``blocked=False`` (default) is the literal double loop of the reference, bit-identical to it.
    ``blocked=True`` evaluates the same P(P+1)/2 dot products as one symmetric matrix product
    (summation order differs at the 1e-16 level only -- but note that the TNC step of the 2D
    bandwidth amplifies even that to ~1e-4 in h for pairs with 0.1 < |corr| <= 0.2).
    """
    norm = np.sum(w)
    if means is None:
        means = weighted_means(X, w)
    n = X.shape[1]
    if blocked:
        D = X - means
        cov = (D * w[:, None]).T.dot(D)
        cov = (cov + cov.T) / 2
        return cov / norm
    diffs = [X[:, i] - means[i] for i in range(n)]
    cov = np.empty((n, n))
    for i, diff in enumerate(diffs):
        wd = diff * w
        for j in range(i, n):
            cov[i, j] = wd.dot(diffs[j])
            cov[j, i] = cov[i, j]
    return cov / norm


def cov_to_corr(cov):
    """chains.py:155-169: divide rows and columns by sqrt(diag), skipping zero diagonals."""
    c = cov.copy()
    for i, di in enumerate(np.sqrt(cov.diagonal())):
        if di:
            c[i, :] /= di
            c[:, i] /= di
    return c


def gelman_rubin_eigenvalues(X, w, chain_offsets, nparam=None):
    """chains.py:1446-1474: var(chain means)/mean(chain var) in the orthogonalised parameters."""
    nparam = nparam or X.shape[1]
    means = weighted_means(X, w)[:nparam]
    nch = len(chain_offsets) - 1
    meanscov = np.zeros((nparam, nparam))
    meancov = np.zeros((nparam, nparam))
    for a, b in zip(chain_offsets[:-1], chain_offsets[1:]):
        Xc, wc = X[a:b], w[a:b]
        mc = weighted_means(Xc, wc)
        d = mc[:nparam] - means
        meanscov += np.outer(d, d)
        meancov += weighted_cov(Xc, wc, mc)[:nparam, :nparam]
    meanscov /= nch - 1
    meancov /= nch
    ev, U = np.linalg.eigh(meancov)
    if np.min(ev) > 0:
        U = U / np.sqrt(ev)
        return np.linalg.eigvalsh(np.dot(U.T, meanscov).dot(U))
    return None


def gelman_rubin(X, w, chain_offsets, nparam=None):
    """chains.py:1476-1486."""
    return np.max(gelman_rubin_eigenvalues(X, w, chain_offsets, nparam))


def weighted_quantiles(x, w, fracs):
    """chains.py:793-838: exact weighted order statistics.

    argsort, cumulative weight in sorted order, ``searchsorted(cumsum, norm*f)`` (side='left'),
    clamped to the last sample.
    """
    order = x.argsort()
    cumsum = np.cumsum(w[order])
    norm = np.sum(w)
    ix = np.searchsorted(cumsum, norm * np.asarray(fracs))
    return x[order[np.minimum(ix, x.shape[0] - 1)]]


def neff_uncorrelated(w):
    """chains.py:500-501: ``(sum w)^2 / sum w^2`` (sampler 'nested'/'uncorrelated')."""
    return np.sum(w) ** 2 / np.dot(w, w)


def auto_convolve(x, n=None, normalize=True):
    """convolve.py:458-478 (``autoConvolve``): result[k] = sum_i x_i x_{i+k} (/ number of terms), evaluated with
    a zero-padded real FFT (the reference packs the same transform through fftpack.rfft + a type-1 DCT)."""
    size = 1
    while size < 2 * x.size:
        size *= 2
    xt = np.fft.rfft(x, size)
    res = np.fft.irfft(xt * np.conj(xt), size)[: (n or x.size)]
    if normalize:
        res = res / np.arange(x.size, x.size - (n or x.size), -1)
    return res


def correlation_length_rows(x, w, mean, var, min_corr=0.05):
    """chains.py:423-466 with weight_units=False: auto-correlation of (x - mean) w up to N//10 lags, summed up to
    the first lag whose value is <= min_corr * corr[0]."""
    d = (x - mean) * w
    corr = auto_convolve(d, n=x.size // 10 + 1, normalize=True) / var
    ix = np.argmin(corr > min_corr * corr[0])
    return corr[0] + 2 * np.sum(corr[1:ix])


def neff_mcmc(x, w, scale, h=0.2, maxoff=None, min_corr=0.05):
    """chains.py:477-574 (``getEffectiveSamplesGaussianKDE``), mcmc branch."""
    norm = np.sum(w)
    mean = w.dot(x) / norm
    var = w.dot((x - mean) ** 2) / norm
    n = x.size
    kernel_std = (scale or np.sqrt(var)) * h
    if maxoff is None:
        maxoff = int(correlation_length_rows(x, w, mean, var) * 1.5) + 4
    maxoff = min(maxoff, n // 10)
    uncorr_len = n // 2
    uncorr_term = 0
    nav = 0
    for k in range(uncorr_len, uncorr_len + 5):
        nav += n - k
        diff2 = (x[:-k] - x[k:]) ** 2 / kernel_std**2
        uncorr_term += np.dot(np.exp(-diff2 / 4) * w[:-k], w[k:])
    uncorr_term /= nav
    corr0 = np.dot(w, w)
    nn = float(n)

    def corr_k(k):
        return np.dot(np.exp(-((x[:-k] - x[k:]) ** 2) / (4 * kernel_std**2)) * w[:-k], w[k:]) - (nn - k) * uncorr_term

    threshold = min_corr * corr0
    c1 = corr_k(1)
    if c1 < threshold:
        N = corr0
    else:
        c2 = corr_k(2)
        if c2 > threshold:
            max_k = maxoff
            while max_k > 10:
                if corr_k(max_k // 3) >= threshold:
                    break
                max_k //= 3
            step_size = 1 if max_k < 20 else max_k // 10
            cum_sum = c1 + c2
            for k in range(3, maxoff + 1, step_size):
                test_val = corr_k(k)
                if test_val < threshold:
                    break
                if k > 3:
                    cum_sum += test_val * step_size
                else:
                    cum_sum += (test_val * step_size) / 2
            N = corr0 + 2 * cum_sum
        else:
            N = corr0 + 2 * c1
    return norm**2 / N


# --------------------------------------------------------------------------------------
# parameter ranges (host-trivial scalar logic once quantiles/min/max exist)
# --------------------------------------------------------------------------------------
def init_param(par: ParamState, x, w, mean, sddev, range_confidence=0.001):
    """mcsamples.py:1427-1484 (``_initParam``) with range_ND_contour off (default ini)."""
    par.err = sddev
    par.mean = mean
    par.param_min = np.min(x)
    par.param_max = np.max(x)
    fr = np.array([range_confidence, 1 - range_confidence] + list(np.linspace(0.1, 0.9, 9)))
    confids = weighted_quantiles(x, w, fr)
    finish_param_ranges(par, confids)
    return par


def finish_param_ranges(par: ParamState, confids):
    """The scalar tail of mcsamples.py:1444-1484, split out so the product's host code can be
    compared step by step (it receives the 11 quantiles from the device)."""
    confids = np.array(confids, dtype=np.float64)
    par.range_min, par.range_max = confids[0:2]
    confids[1:-1] = confids[2:]
    confids[0] = par.param_min
    confids[-1] = par.param_max
    diffs = confids[4:] - confids[:-4]
    scale = np.min(diffs) / 1.049
    if np.all(diffs > par.err * 1.049) and np.all(diffs < scale * 1.5):
        par.sigma_range = scale
    else:
        par.sigma_range = min(par.err, scale)
    smooth_1D = par.sigma_range * 0.4
    par.has_limits_bot = par.limmin is not None
    par.has_limits_top = par.limmax is not None
    if par.has_limits_bot:
        if par.range_min - par.limmin > 2 * smooth_1D and par.param_min - par.limmin > smooth_1D:
            par.has_limits_bot = False
        else:
            par.range_min = par.limmin
    if par.has_limits_top:
        if par.limmax - par.range_max > 2 * smooth_1D and par.limmax - par.param_max > smooth_1D:
            par.has_limits_top = False
        else:
            par.range_max = par.limmax
    if not par.has_limits_bot:
        par.range_min -= smooth_1D * 2
    if not par.has_limits_top:
        par.range_max += smooth_1D * 2
    par.has_limits = par.has_limits_top or par.has_limits_bot
    return par


def bin_geometry(par: ParamState, num_fine_bins, borderfrac=0.1):
    """mcsamples.py:1486-1496: grid geometry (first and last bins are half width)."""
    border = (par.range_max - par.range_min) * borderfrac
    binmin = min(par.param_min, par.range_min)
    if not par.has_limits_bot:
        binmin -= border
    binmax = max(par.param_max, par.range_max)
    if not par.has_limits_top:
        binmax += border
    fine_width = (binmax - binmin) / (num_fine_bins - 1)
    return binmin, binmax, fine_width


def bin_indices(x, binmin, fine_width):
    """mcsamples.py:1497: round-half-up of a non-negative value by truncation."""
    return ((x - binmin) / fine_width + 0.5).astype(int)


def kde_bin_samples(x, range_min=None, range_max=None, nbins=2046, edge_fac=0.1):
    """kde_bandwidth.py:76-87: truncating binning used for the sheared 2D re-binning."""
    mx = np.max(x)
    mn = np.min(x)
    delta = mx - mn
    if range_min is None:
        range_min = mn - delta * edge_fac
    if range_max is None:
        range_max = mx + delta * edge_fac
    R = range_max - range_min
    dx = R / (nbins - 1)
    return ((x - range_min) / dx).astype(int), R


# --------------------------------------------------------------------------------------
# 1D bandwidth: Botev improved Sheather-Jones on the binned data
# --------------------------------------------------------------------------------------
_ROOTPI = np.sqrt(np.pi)
_PI2 = np.pi**2
_LMAX = 7
# kde_bandwidth.py:50-56
_ISJ_CONSTS = np.array(
    [
        (1 + 0.5 ** (j + 0.5)) / 3 * np.prod(np.arange(1, 2 * j, 2)) / (_ROOTPI / np.sqrt(2.0))
        for j in range(_LMAX - 1, 1, -1)
    ]
)


def isj_fixed_point(h, N, I, logI, a2):
    """kde_bandwidth.py:59-73 (``_bandwidth_fixed_point``)."""
    if h <= 0:
        return h - 1
    f = 2 * np.pi ** (2 * _LMAX) * np.dot(a2, np.exp(_LMAX * logI - I * (_PI2 * h**2)))
    for j, const in zip(range(_LMAX - 1, 1, -1), _ISJ_CONSTS):
        t_j = (const / N / f) ** (2 / (3.0 + 2 * j))
        f = 2 * np.pi ** (2 * j) * np.dot(a2, np.exp(j * logI - I * (_PI2 * t_j)))
        if not f:
            raise Exception("zero f in _bandwidth_fixed_point (non-convergence)")
    return h - (2 * N * _ROOTPI * f) ** (-1.0 / 5)


def isj_bandwidth_binned(data, neff):
    """kde_bandwidth.py:102-135 (``gaussian_kde_bandwidth_binned``): bandwidth as a fraction of
    the binned range, or None on failure."""
    I = np.arange(1, data.size) ** 2
    logI = np.log(I)
    a = fftpack.dct(data / np.sum(data))
    a2 = (a[1:] / 2) ** 2
    try:
        n_scaling = neff ** (-1.0 / 5)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h0 = 0.53 * n_scaling
            h = fsolve(isj_fixed_point, h0, (neff, I, logI, a2), xtol=h0 / 20, factor=1)[0]
        if h < 0.019 * n_scaling:
            try:
                h = brentq(isj_fixed_point, 0.019 * n_scaling, 0.5, (neff, I, logI, a2), xtol=h / 20)
            except Exception:
                pass
        return h
    except Exception as e:  # noqa
        log.warning("1D auto bandwidth failed. Using fallback: %s", e)
        return None


# --------------------------------------------------------------------------------------
# 2D bandwidth: plug-in psi functionals on the DCT / FFT of the binned data
# --------------------------------------------------------------------------------------
_K2D = np.array(
    [1 / np.sqrt(2 * np.pi)] + [(-1) ** j * np.prod(np.arange(1, 2 * j, 2)) / np.sqrt(2 * np.pi) for j in range(1, 5)]
)
_KODD = np.array([1] + [np.prod(np.arange(1, 2 * j, 2)) / 2.0 ** (j + 1) / np.sqrt(np.pi) for j in range(1, 9)])


class BandwidthOptimizer2D:
    """kde_bandwidth.py:146-309 (``KernelOptimizer2D``).

    The reference evaluates the ``func2d`` / ``func2d_odd`` recursions without memoisation (45
    psi evaluations per fixed-point evaluation, many repeated); here each distinct (s, t) node is
    evaluated once.  The arithmetic of every node is identical, so the results are bit-identical
    to the unmemoised recursion.
    """

    def __init__(self, data, neff, correlation, do_correlation=True, fallback_t=None):
        size = data.shape[0]
        if size != data.shape[1]:
            raise ValueError("square arrays only")
        p = data / np.sum(data)
        self.a2 = fftpack.dct(fftpack.dct(p, axis=0), axis=1)[1:, 1:] ** 2  # convolve.py:565-566
        self.I = np.arange(1, size, dtype=np.float64) ** 2
        self.logI = np.log(self.I)
        self.do_correlation = do_correlation
        if do_correlation:
            # kept complex (imaginary part exactly 0) so that the BLAS call below is the same
            # zgemv the reference issues -- the TNC step downstream amplifies ulp-level changes
            self.aFFT = np.fft.fft2(p)
            self.aFFT *= np.conj(self.aFFT)
            n = size
            self.freq = np.fft.fftfreq(n, d=1.0 / n)
        self.N = neff
        self.corr = correlation
        self.n_brent_evals = 0
        self.used_fallback_t = False
        try:
            self.t_star = brentq(self._fixed_point, 0, 0.1, xtol=0.001**2)
            if fallback_t and self.t_star > 0.01 and self.t_star > 2 * fallback_t:
                self.t_star = fallback_t
                self.used_fallback_t = True
        except Exception:
            if fallback_t is not None:
                self.t_star = fallback_t
                self.used_fallback_t = True
            else:
                raise

    # -- even functionals -----------------------------------------------------------
    def psi(self, s0, s1, time):
        """kde_bandwidth.py:182-186."""
        w = -self.I * (_PI2 * time)
        wx = np.exp(w + self.logI * s0)
        wy = np.exp(w + self.logI * s1)
        ssum = s0 + s1
        return (-1) ** ssum * wy.dot(self.a2).dot(wx.T) * np.pi ** (2 * ssum) / 4

    def func2d_table(self, t, min_sum):
        """kde_bandwidth.py:188-196 for every s with min_sum <= |s| <= 5, filled top-down."""
        tab = {}
        for ssum in range(5, min_sum - 1, -1):
            for s0 in range(ssum + 1):
                s1 = ssum - s0
                if ssum > 4:
                    tab[(s0, s1)] = self.psi(s0, s1, t)
                else:
                    sum_func = tab[(s0 + 1, s1)] + tab[(s0, s1 + 1)]
                    const = (1 + 0.5 ** (ssum + 1)) / 3
                    time = (-2 * const * _K2D[s0] * _K2D[s1] / self.N / sum_func) ** (1.0 / (2 + ssum))
                    tab[(s0, s1)] = self.psi(s0, s1, time)
        return tab

    def _fixed_point(self, t):
        """kde_bandwidth.py:177-180."""
        self.n_brent_evals += 1
        tab = self.func2d_table(t, 2)
        sum_func = tab[(0, 2)] + tab[(2, 0)] + 2 * tab[(1, 1)]
        time = (2 * np.pi * self.N * sum_func) ** (-1.0 / 3)
        return (t - time) / time

    # -- odd functionals ------------------------------------------------------------
    def psi_odd(self, s0, s1, time):
        """kde_bandwidth.py:209-214."""
        f = self.freq
        w = np.exp(-(f**2) * (4 * _PI2 * time))
        wx = w * f**s0
        wy = w * f**s1
        return wy.dot(self.aFFT).real.dot(wx.T) * (2 * np.pi) ** (s0 + s1)

    def func2d_odd_table(self, t, p00):
        """kde_bandwidth.py:198-207 for the odd-odd nodes under [1,3] and [3,1]."""
        tab = {}
        for ssum in (10, 8, 6, 4):
            for s0 in range(1, ssum, 2):
                s1 = ssum - s0
                if ssum > 8:
                    tab[(s0, s1)] = self.psi_odd(s0, s1, t)
                else:
                    sum_func = tab[(s0 + 2, s1)] + tab[(s0, s1 + 2)]
                    const = 8 * (1 - 2.0 ** (-ssum - 1)) / 3.0
                    time = (const * p00 * _KODD[s0] * _KODD[s1] / self.N**2 / sum_func**2) ** (1.0 / (3 + ssum))
                    tab[(s0, s1)] = self.psi_odd(s0, s1, time)
        return tab

    def amise(self, cov, corr=None):
        """kde_bandwidth.py:216-232."""
        hx, hy = cov[0], cov[1]
        c = corr if corr is not None else cov[2]
        p = self.p
        var = 1.0 / (4 * np.pi * hx * hy * np.sqrt(1 - c**2) * self.N)
        bias = 0.25 * (
            hx**4 * p[4, 0]
            + hy**4 * p[0, 4]
            + 2 * hx**2 * hy**2 * p[2, 2] * (2 * c**2 + 1)
            + 4 * c * hx * hy * (hx**2 * p[3, 1] + hy**2 * p[1, 3])
        )
        if bias < 0:
            raise Exception("bias not positive definite")
        return var + bias

    def get_h(self):
        """kde_bandwidth.py:234-306."""
        tab = self.func2d_table(self.t_star, 0 if self.do_correlation else 2)
        p_02, p_20, p_11 = tab[(0, 2)], tab[(2, 0)], tab[(1, 1)]
        h_x = (p_02 ** 0.75 / (4 * np.pi * self.N * p_20**0.75 * (p_11 + np.sqrt(p_20 * p_02)))) ** (1.0 / 6)
        h_y = (p_20 ** 0.75 / (4 * np.pi * self.N * p_02**0.75 * (p_11 + np.sqrt(p_20 * p_02)))) ** (1.0 / 6)
        self.h_closed = (h_x, h_y)
        corr = 0
        if not self.do_correlation:
            return h_x, h_y, corr
        p = np.zeros((5, 5))
        p[0, 4], p[4, 0], p[2, 2] = p_02, p_20, p_11
        p[0, 0] = tab[(0, 0)]
        odd = self.func2d_odd_table(self.t_star, p[0, 0])
        p[1, 3], p[3, 1] = odd[(1, 3)], odd[(3, 1)]
        self.p = p
        AMISE = self.amise(np.array([h_x, h_y, 0]))
        if self.corr:
            try:
                res = minimize(
                    self.amise,
                    np.array([h_x, h_y]) / np.sqrt(1 - abs(self.corr)),
                    (self.corr,),
                    method="TNC",
                    bounds=[(0.001, 0.3), (0.001, 0.3)],
                )
                if res.success:
                    A2 = self.amise(res.x, self.corr)
                    if A2 < AMISE:
                        h_x, h_y = res.x
                        corr = self.corr
                        AMISE = A2
            except Exception:  # noqa
                pass
        try:
            res = minimize(
                self.amise,
                np.array([h_x, h_y, self.corr]),
                (None,),
                method="TNC",
                bounds=[(0.001, 0.3), (0.001, 0.3), (-0.99, 0.99)],
            )
            if res.success:
                A3 = self.amise(res.x)
                if A3 < AMISE * 0.9:
                    h_x, h_y, corr = res.x
        except Exception:  # noqa
            pass
        return h_x, h_y, corr


# --------------------------------------------------------------------------------------
# convolution helpers (algorithm free to choose: SURVEY s8c, differences <= 4e-15)
# --------------------------------------------------------------------------------------
def conv1d(x, k, mode):
    """convolve.py:196-202: direct ``np.convolve`` unless both operands exceed 1000 taps."""
    if min(x.shape[0], k.shape[0]) > 1000:
        return fftconvolve(x, k, mode)
    return np.convolve(x, k, mode)


def conv2d(x, k, mode):
    """convolve.py:205-212, 405-436: FFT linear convolution, modes 'same' / 'valid'; periodic modes :215-323."""
    if mode in ("periodic_both", "periodic_x", "periodic_y"):
        return conv2d_periodic(x, k, periodic_x=mode != "periodic_y", periodic_y=mode != "periodic_x")
    return fftconvolve(x, k, mode)


def conv1d_periodic(x, k):
    """convolve.py:326-367: the last bin is folded onto the first, circular convolution of length n-1 with the
    centred kernel, result extended by its first element."""
    xc = x[:-1].copy()
    xc[0] += x[-1]
    N, M = xc.shape[0], k.shape[0]
    hpad = np.zeros(N)
    hpad[:M] = k
    hpad = np.roll(hpad, -(M // 2))
    res = np.fft.irfft(np.fft.rfft(xc) * np.fft.rfft(hpad), n=N)
    return np.append(res, res[0])


def conv2d_periodic(x, k, periodic_x=True, periodic_y=True):
    """convolve.py:215-323.  NB the transform has the size of the folded array in BOTH axes, so the convolution
    is circular along a non-periodic axis as well (as in the reference)."""
    ny, nx = x.shape
    ky, kx = k.shape
    if periodic_x and periodic_y:
        xc = x[:-1, :-1].copy()
        xc[0, :] += x[-1, :-1]
        xc[:, 0] += x[:-1, -1]
        xc[0, 0] += x[-1, -1]
    elif periodic_x:
        xc = x[:, :-1].copy()
        xc[:, 0] += x[:, -1]
    else:
        xc = x[:-1, :].copy()
        xc[0, :] += x[-1, :]
    Ny, Nx = xc.shape
    hpad = np.zeros((Ny, Nx))
    hpad[:ky, :kx] = k
    hpad = np.roll(np.roll(hpad, -(ky // 2), axis=0), -(kx // 2), axis=1)
    res = np.fft.irfftn(np.fft.rfftn(xc) * np.fft.rfftn(hpad), (Ny, Nx), axes=(0, 1))
    out = np.empty((ny, nx))
    out[:Ny, :Nx] = res
    if periodic_y:
        out[-1, :Nx] = res[0, :]
    if periodic_x:
        out[:Ny, -1] = res[:, 0]
    if periodic_x and periodic_y:
        out[-1, -1] = res[0, 0]
    return out


# --------------------------------------------------------------------------------------
# the sample container mirroring the MCSamples surface of the hot path
# --------------------------------------------------------------------------------------
class OracleSamples:
    """Restates the MCSamples/Chains/WeightedSamples hot-path surface."""

    def __init__(self, samples, weights=None, names=None, ranges=None, sampler="uncorrelated", settings=None,
                 chain_offsets=None, loglikes=None):
        if isinstance(samples, (list, tuple)):
            # chains.py:1488-1503 (makeSingle)
            chain_offsets = np.cumsum([0] + [s.shape[0] for s in samples])
            if weights is not None:
                weights = np.hstack(list(weights))
            if loglikes is not None:
                loglikes = np.hstack(list(loglikes))
            samples = np.vstack(list(samples))
        self.loglikes = None if loglikes is None else np.asarray(loglikes, dtype=np.float64)
        self.samples = np.asarray(samples, dtype=np.float64)
        self.numrows, self.n = self.samples.shape
        self.weights = np.ones(self.numrows) if weights is None else np.asarray(weights, dtype=np.float64)
        self.chain_offsets = None if chain_offsets is None else np.asarray(chain_offsets, dtype=np.int64)
        self.names = list(names) if names is not None else ["param%d" % (i + 1) for i in range(self.n)]
        self.index = {n: i for i, n in enumerate(self.names)}
        self.sampler = sampler
        if sampler not in ("uncorrelated", "nested", "mcmc"):
            raise OracleError("unknown sampler")
        self.settings = dict(DEFAULT_SETTINGS)
        if settings:
            self.settings.update(settings)
        self.ranges = dict(ranges or {})
        self.pars = [ParamState(name=n) for n in self.names]
        self.raise_on_bandwidth_errors = False
        self.update_base_statistics()

    # chains.py:1340-1352 + mcsamples.py:552-576
    def update_base_statistics(self):
        self.norm = np.sum(self.weights)
        self.means = weighted_means(self.samples, self.weights)
        # chains.py:380-383
        self.mean_loglike = None if self.loglikes is None else self.weights.dot(self.loglikes) / self.norm
        self.vars = weighted_vars(self.samples, self.weights, self.means)
        self.sddev = np.sqrt(self.vars)
        self.mean_mult = self.norm / self.numrows
        self.max_mult = np.max(self.weights)
        self.fullcov = None
        self.corrmat = None
        for par in self.pars:
            rg = self.ranges.get(par.name, (None, None))
            lo, hi = rg[0], rg[1]
            par.periodic = len(rg) > 2 and (rg[2] is True or (isinstance(rg[2], str) and rg[2].upper() in ("T", "TRUE", "PERIODIC")))
            par.limmin, par.limmax = lo, hi
            par.has_limits_bot = lo is not None
            par.has_limits_top = hi is not None
            par.N_eff_kde = None
        self.density1D = {}

    def _num(self, j):
        return self.index[j] if isinstance(j, str) else int(j)

    def get_means(self):
        return self.means

    def get_vars(self):
        return self.vars

    def get_cov(self, nparam=None, pars=None):
        if self.fullcov is None:
            self.fullcov = weighted_cov(self.samples, self.weights, self.means)
        if pars is not None:
            return self.fullcov[np.ix_(pars, pars)]
        return self.fullcov[:nparam, :nparam]

    def get_correlation_matrix(self):
        if self.corrmat is None:
            self.corrmat = cov_to_corr(self.get_cov())
        return self.corrmat

    def get_gelman_rubin(self, nparam=None):
        return gelman_rubin(self.samples, self.weights, self.chain_offsets, nparam)

    # ---- raw ND densities and the MeanVar / split convergence tests ------------------------------
    def raw_nd_density(self, js, meanlikes=False, maxlikes=False, **kwargs):
        """getRawNDDensityGridData (mcsamples.py:2098-2197): returns (xs, P max-normalised, likes | None, maxlikes | None)."""
        jv = [self._num(j) for j in js]
        parv = [self.init_param_ranges(j) for j in jv]
        ndim = len(jv)
        nb = int(kwargs.get("num_bins_ND", self.settings.get("num_bins_ND", 12)))
        bco = kwargs.get("boundary_correction_order", self.settings["boundary_correction_order"])
        ixv, xminv, xmaxv = [], [], []
        for j, par in zip(jv, parv):
            lo, hi, fw = bin_geometry(par, nb)  # mcsamples.py:1486-1496
            ixv.append(bin_indices(self.samples[:, j], lo, fw))
            xminv.append(lo)
            xmaxv.append(hi)
        flat = np.zeros(self.numrows, dtype=np.int64)  # _flattenValues (mcsamples.py:2034-2046): axis 0 fastest
        stride = 1
        for ix in ixv:
            flat += ix * stride
            stride *= nb
        shape = tuple([nb] * ndim)
        bins = np.bincount(flat, weights=self.weights, minlength=nb ** ndim).reshape(shape)  # :2077-2079
        if any(p.has_limits_bot or p.has_limits_top for p in parv) and bco >= 0:
            mask = np.ones(shape)  # _setRawEdgeMaskND, :2012-2032
            for ax, par in enumerate(parv[::-1]):
                sl = [slice(None)] * ndim
                if par.has_limits_bot:
                    sl[ax] = 0
                    mask[tuple(sl)] /= 2
                if par.has_limits_top:
                    sl[ax] = nb - 1
                    mask[tuple(sl)] /= 2
            bins = bins / mask
        P = bins / np.max(bins)
        likes = mlk = None
        if meanlikes:  # :2155-2159, 2200-2201
            lw = self.weights * np.exp(self.mean_loglike - self.loglikes)
            likes = np.bincount(flat, weights=lw, minlength=nb ** ndim).reshape(shape)
            likes = likes / np.max(likes)
        if maxlikes:  # :2163-2169
            bestfit = np.max(-self.loglikes)
            mlk = np.zeros(nb ** ndim)
            np.maximum.at(mlk, flat, np.exp(-bestfit - self.loglikes))
            mlk = mlk.reshape(shape)
        xs = [np.linspace(xminv[i], xmaxv[i], nb) for i in range(ndim)]
        return xs, P, likes, mlk

    def fraction_indices(self, n):
        """getFractionIndices (mcsamples.py:668-680)."""
        cumsum = np.cumsum(self.weights)
        return np.append(np.searchsorted(cumsum, np.linspace(0, 1, n, endpoint=False) * self.norm), self.numrows)

    def mean_var_test(self):
        """'MeanVar' of getConvergeTests (mcsamples.py:964-989)."""
        offs = self.chain_offsets
        nch = len(offs) - 1
        between = np.zeros(self.n)
        within = np.zeros(self.n)
        for a, b in zip(offs[:-1], offs[1:]):
            w = self.weights[a:b]
            cm = w.dot(self.samples[a:b]) / np.sum(w)
            between += (cm - self.means) ** 2
            for j in range(self.n):
                within[j] += np.dot(w, (self.samples[a:b, j] - cm[j]) ** 2)
        return np.sqrt(between / (nch - 1) / (within / self.norm))

    def split_tests(self, test_confidence=0.95, max_split_tests=4):
        """'SplitTest' of getConvergeTests (mcsamples.py:1003-1034) -> (nparam, max_split_tests - 1, 2)."""
        limits = np.array([1 - (1 - test_confidence) / 2, (1 - test_confidence) / 2])
        out = np.zeros((self.n, max_split_tests - 1, 2))
        fracs = [self.fraction_indices(i + 2) for i in range(max_split_tests - 1)]
        for j in range(self.n):
            confids = weighted_quantiles(self.samples[:, j], self.weights, limits)
            for ix, frac in enumerate(fracs):
                for f1, f2 in zip(frac[:-1], frac[1:]):
                    out[j, ix] += (weighted_quantiles(self.samples[f1:f2, j], self.weights[f1:f2], limits) - confids) ** 2
                out[j, ix] = np.sqrt(out[j, ix] / (2 + ix)) / self.sddev[j]
        return out

    def init_param_ranges(self, j):
        """mcsamples.py:1421-1425; resets the limit flags from the hard ranges as _initLimits does
        once per updateBaseStatistics (the proximity test may clear them, :1463-1474)."""
        j = self._num(j)
        par = self.pars[j]
        par.has_limits_bot = par.limmin is not None
        par.has_limits_top = par.limmax is not None
        return init_param(par, self.samples[:, j], self.weights, self.means[j], self.sddev[j],
                          self.settings["range_confidence"])

    def _neff(self, par):
        """mcsamples.py:1230-1235 with chains.py:500-501."""
        if par.N_eff_kde is None:
            if self.sampler == "mcmc":
                j = self.index[par.name]
                par.N_eff_kde = neff_mcmc(self.samples[:, j], self.weights, par.sigma_range)
            else:
                par.N_eff_kde = neff_uncorrelated(self.weights)
        return par.N_eff_kde

    # ---- 1D ---------------------------------------------------------------------------
    def auto_bandwidth_1d(self, bins, par, mult_bias_correction_order, kernel_order):
        """mcsamples.py:1237-1283."""
        N_eff = self._neff(par)
        h = isj_bandwidth_binned(bins, N_eff)
        bin_range = max(par.param_max, par.range_max) - min(par.param_min, par.range_min)
        if h is None or h < 0.01 * N_eff ** (-1.0 / 5) * (par.range_max - par.range_min) / bin_range:
            hnew = 1.06 * par.sigma_range * N_eff ** (-1.0 / 5) / bin_range
            msg = "auto bandwidth for %s very small or failed (h=%s,N_eff=%s). Using fallback (h=%s)" % (
                par.name, h, N_eff, hnew)
            if self.raise_on_bandwidth_errors:
                raise OracleBandwidthError(msg)
            log.warning(msg)
            h = hnew
        par.kde_h = h
        m = mult_bias_correction_order
        if kernel_order > 1:
            m = max(m, 1)
        if m:
            return h * N_eff ** (1.0 / 5 - 1.0 / (4 * m + 5))
        return h

    def density_1d(self, j, meanlikes=False, **kwargs):
        """mcsamples.py:1517-1686 (``get1DDensityGridData``)."""
        j = self._num(j)
        par = self.init_param_ranges(j)
        s = self.settings
        num_bins = kwargs.get("num_bins", s["num_bins"])
        smooth_scale_1D = kwargs.get("smooth_scale_1D", s["smooth_scale_1D"])
        boundary_correction_order = kwargs.get("boundary_correction_order", s["boundary_correction_order"])
        mult_bias_correction_order = kwargs.get("mult_bias_correction_order", s["mult_bias_correction_order"])
        fine_bins = kwargs.get("fine_bins", s["fine_bins"])

        paramrange = par.range_max - par.range_min
        if paramrange <= 0:
            raise OracleError("Parameter range is <= 0: " + par.name)
        width = paramrange / (num_bins - 1)
        binmin, binmax, fine_width = bin_geometry(par, fine_bins)
        ix = bin_indices(self.samples[:, j], binmin, fine_width)
        bins = np.bincount(ix, weights=self.weights, minlength=fine_bins)
        if meanlikes:  # mcsamples.py:1556-1561
            if s.get("shade_likes_is_mean_loglikes", False):
                lw = self.weights * self.loglikes
            else:
                lw = self.weights * np.exp(self.mean_loglike - self.loglikes)
            finebinlikes = np.bincount(ix, weights=lw, minlength=fine_bins)

        if smooth_scale_1D <= 0:
            bandwidth = self.auto_bandwidth_1d(bins, par, mult_bias_correction_order, boundary_correction_order) * (
                binmax - binmin)
            bandwidth = min(bandwidth, paramrange / 4)
            smooth_1D = bandwidth * abs(smooth_scale_1D) / fine_width
        elif smooth_scale_1D < 1.0:
            smooth_1D = smooth_scale_1D * par.err / fine_width
        else:
            smooth_1D = smooth_scale_1D * width / fine_width
        if smooth_1D < 2:
            log.warning("fine_bins not large enough to well sample smoothing scale - " + par.name)
        smooth_1D = min(max(1.0, smooth_1D), fine_bins // 2)
        winw = min(int(round(2.5 * smooth_1D)), ((fine_bins - 1) if par.periodic else fine_bins) // 2 - 2)
        kx = np.arange(-winw, winw + 1)
        Win = np.exp(-((kx / smooth_1D) ** 2) / 2.0)
        Win = Win / np.sum(Win)  # mcsamples.py:129-135

        def cconv(a):  # mcsamples.py:1592-1593: convolution mode follows the parameter
            return conv1d_periodic(a, Win) if par.periodic else conv1d(a, Win, "same")

        P = cconv(bins)
        if meanlikes:
            rawbins = P.copy()  # mcsamples.py:1597-1598
        if par.has_limits and not par.periodic and boundary_correction_order >= 0:
            # mcsamples.py:1600-1637
            prior_mask = np.ones(fine_bins + 2 * winw)
            if par.has_limits_bot:
                prior_mask[winw] = 0.5
                prior_mask[:winw] = 0
            if par.has_limits_top:
                prior_mask[-(winw + 1)] = 0.5
                prior_mask[-winw:] = 0
            a0 = conv1d(prior_mask, Win, "valid")
            sel = np.nonzero(a0 * P)
            a0 = a0[sel]
            normed = P[sel] / a0
            if boundary_correction_order == 0:
                P[sel] = normed
            elif boundary_correction_order <= 2:
                xWin = Win * kx
                a1 = conv1d(prior_mask, xWin, "valid")[sel]
                a2 = conv1d(prior_mask, xWin * kx, "valid")[sel]
                xP = conv1d(bins, xWin, "same")[sel]
                if boundary_correction_order == 1:
                    corrected = (P[sel] * a2 - xP * a1) / (a0 * a2 - a1**2)
                else:
                    a3 = conv1d(prior_mask, xWin * kx**2, "valid")[sel]
                    a4 = conv1d(prior_mask, xWin * kx**3, "valid")[sel]
                    x2P = conv1d(bins, xWin * kx, "same")[sel]
                    denom = a4 * a2 * a0 - a4 * a1**2 - a2**3 - a3**2 * a0 + 2 * a1 * a2 * a3
                    A = a4 * a2 - a3**2
                    B = a2 * a3 - a4 * a1
                    C = a3 * a1 - a2**2
                    corrected = (P[sel] * A + xP * B + x2P * C) / denom
                P[sel] = normed * np.exp(np.minimum(corrected / normed, 4) - 1)
            else:
                raise OracleSettingError("Unknown boundary_correction_order (expected 0, 1, 2)")
        elif not par.periodic and boundary_correction_order == 2:
            # mcsamples.py:1638-1647
            xWin2 = Win * kx**2
            x2P = conv1d(bins, xWin2, "same")
            a2 = np.sum(xWin2)
            a4 = np.dot(xWin2, kx**2)
            corrected = (P * a4 - a2 * x2P) / (a4 - a2**2)
            sel = P > 0
            P[sel] *= np.exp(np.minimum(corrected[sel] / P[sel], 2) - 1)

        if mult_bias_correction_order:
            # mcsamples.py:1649-1666
            prior_mask = np.ones(fine_bins)
            if par.has_limits_bot:
                prior_mask[0] *= 0.5
            if par.has_limits_top:
                prior_mask[-1] *= 0.5
            a0 = conv1d(prior_mask, Win, "same")
            for _ in range(mult_bias_correction_order):
                prob1 = P.copy()
                prob1[prob1 == 0] = 1
                fine = bins / prob1
                conv = cconv(fine)
                P = P * conv
                if not par.periodic:
                    P /= a0
        mx = np.max(P)
        if mx == 0:
            raise OracleDensityError("no samples in bin")
        P = P / mx  # densities.py:71-92, by='max'
        x = np.linspace(binmin, binmax, fine_bins)
        likes = None
        if meanlikes:  # mcsamples.py:1672-1682
            sel = P > 0
            finebinlikes[sel] /= P[sel]
            binlikes = cconv(finebinlikes)
            binlikes[sel] *= P[sel] / rawbins[sel]
            if s.get("shade_likes_is_mean_loglikes", False):
                maxbin = np.min(binlikes)
                binlikes = np.where((binlikes - maxbin) < 30, np.exp(-(binlikes - maxbin)), 0)
                binlikes[rawbins == 0] = 0
            binlikes /= np.max(binlikes)
            likes = binlikes
        return Grid1D(x, P, (par.range_min, par.range_max), h=par.kde_h, winw=winw, smooth_bins=smooth_1D, likes=likes)

    # ---- 2D ---------------------------------------------------------------------------
    def _hist2d(self, ixs, iys, xsize, ysize):
        """mcsamples.py:1724-1728: arrays are indexed [y, x]."""
        flat = ixs + iys * xsize
        return np.bincount(flat, weights=self.weights, minlength=xsize * ysize).reshape((ysize, xsize))

    def auto_bandwidth_2d(self, bins, parx, pary, jx, jy, corr, rangex, rangey, base_fine_bins_2D,
                          mult_bias_correction_order, min_corr=0.2):
        """mcsamples.py:1285-1419."""
        max_corr_2D = self.settings["max_corr_2D"]
        N_eff = min(self._neff(parx), self._neff(pary))
        has_limits = parx.has_limits or pary.has_limits
        do_correlated = not parx.has_limits or not pary.has_limits
        info = {"branch": None}

        def fallback_widths(ex):
            msg = "2D kernel density bandwidth optimizer failed for %s, %s. Using fallback width: %s" % (
                parx.name, pary.name, ex)
            if self.raise_on_bandwidth_errors:
                raise OracleBandwidthError(msg)
            log.warning(msg)
            info["fallback"] = True
            return (parx.sigma_range / N_eff ** (1.0 / 6), pary.sigma_range / N_eff ** (1.0 / 6),
                    max(min(corr, max_corr_2D), -max_corr_2D))

        if min_corr < abs(corr) <= max_corr_2D and do_correlated:
            info["branch"] = "shear"
            i, j = jx, jy
            imax, imin = None, None
            if parx.has_limits_bot:
                imin = parx.range_min
            if parx.has_limits_top:
                imax = parx.range_max
            if pary.has_limits:
                i, j = j, i
                if pary.has_limits_bot:
                    imin = pary.range_min
                if pary.has_limits_top:
                    imax = pary.range_max
            cov = self.get_cov(pars=[i, j])
            S = np.linalg.cholesky(cov)
            ichol = np.linalg.inv(S)
            S = S * ichol[0, 0]
            r = ichol[1, :] / ichol[0, 0]
            p1 = self.samples[:, i]
            p2 = r[0] * self.samples[:, i] + r[1] * self.samples[:, j]
            bin1, r1 = kde_bin_samples(p1, nbins=base_fine_bins_2D, range_min=imin, range_max=imax)
            bin2, r2 = kde_bin_samples(p2, nbins=base_fine_bins_2D)
            rotbins = self._hist2d(bin1, bin2, base_fine_bins_2D, base_fine_bins_2D)
            try:
                opt = BandwidthOptimizer2D(rotbins, N_eff, 0, do_correlation=not has_limits)
                hx, hy, c = opt.get_h()
                info["opt"] = opt
                hx *= r1
                hy *= r2
                kernelC = S.dot(np.array([[hx**2, hx * hy * c], [hx * hy * c, hy**2]])).dot(S.T)
                hx, hy, c = (np.sqrt(kernelC[0, 0]), np.sqrt(kernelC[1, 1]),
                             kernelC[0, 1] / np.sqrt(kernelC[0, 0] * kernelC[1, 1]))
                if pary.has_limits:
                    hx, hy = hy, hx
            except ValueError as e:
                hx, hy, c = fallback_widths(e)
        elif abs(corr) > max_corr_2D or not do_correlated and corr > 0.8:
            info["branch"] = "rule"
            c = max(min(corr, max_corr_2D), -max_corr_2D)
            hx = parx.sigma_range / N_eff ** (1.0 / 6)
            hy = pary.sigma_range / N_eff ** (1.0 / 6)
        else:
            info["branch"] = "plain"
            try:
                opt = BandwidthOptimizer2D(
                    bins, N_eff, corr, do_correlation=not has_limits,
                    fallback_t=(min(pary.sigma_range / rangey, parx.sigma_range / rangex) / N_eff ** (1.0 / 6)) ** 2)
                hx, hy, c = opt.get_h()
                info["opt"] = opt
                hx *= rangex
                hy *= rangey
            except ValueError as e:
                hx, hy, c = fallback_widths(e)
        if mult_bias_correction_order:
            scale = 1.1 * N_eff ** (1.0 / 6 - 1.0 / (2 + 4 * (1 + mult_bias_correction_order)))
            hx *= scale
            hy *= scale
        return hx, hy, c, info

    def density_2d(self, j, j2, meanlikes=False, mask_function=None, **kwargs):
        """mcsamples.py:1748-2010 (``get2DDensityGridData``; ``likes`` as in the get_density=False return)."""
        j, j2 = self._num(j), self._num(j2)
        parx = self.init_param_ranges(j)
        pary = self.init_param_ranges(j2)
        s = self.settings
        base_fine_bins_2D = kwargs.get("fine_bins_2D", s["fine_bins_2D"])
        boundary_correction_order = kwargs.get("boundary_correction_order", s["boundary_correction_order"])
        mult_bias_correction_order = kwargs.get("mult_bias_correction_order", s["mult_bias_correction_order"])
        smooth_scale_2D = float(kwargs.get("smooth_scale_2D", s["smooth_scale_2D"]))
        max_corr_2D = s["max_corr_2D"]
        has_prior = parx.has_limits or pary.has_limits or mask_function is not None  # mcsamples.py:1794

        corr = self.get_correlation_matrix()[j2][j]
        actual_corr = corr
        if abs(abs(corr) - 1.0) <= 1e-8:
            corr = np.sign(corr) * max_corr_2D
        if abs(max_corr_2D) > 1:
            raise OracleSettingError("max_corr_2D cannot be >=1")
        if abs(corr) < 0.1:
            corr = 0.0
        angle_scale = max(0.2, np.sqrt(1 - min(max_corr_2D, abs(corr)) ** 2))
        nbin2D = int(round(s["num_bins_2D"] / angle_scale))
        fine_bins_2D = base_fine_bins_2D
        if corr:
            scaled = 192 * int(3 / angle_scale) // 3
            if base_fine_bins_2D < scaled and int(1 / angle_scale) > 1:
                fine_bins_2D = scaled

        xbinmin, xbinmax, finewidthx = bin_geometry(parx, fine_bins_2D)
        ybinmin, ybinmax, finewidthy = bin_geometry(pary, fine_bins_2D)
        ixs = bin_indices(self.samples[:, j], xbinmin, finewidthx)
        iys = bin_indices(self.samples[:, j2], ybinmin, finewidthy)
        xsize = ysize = fine_bins_2D
        histbins = self._hist2d(ixs, iys, xsize, ysize)
        if meanlikes:  # mcsamples.py:1829-1831
            likeweights = self.weights * np.exp(self.mean_loglike - self.loglikes)
            finebinlikes = np.bincount(ixs + iys * xsize, weights=likeweights, minlength=xsize * ysize).reshape(
                (ysize, xsize))
        info = {}
        if smooth_scale_2D < 0:
            rx, ry, corr, info = self.auto_bandwidth_2d(
                histbins, parx, pary, j, j2, actual_corr, xbinmax - xbinmin, ybinmax - ybinmin,
                base_fine_bins_2D, mult_bias_correction_order)
            rx = rx * abs(smooth_scale_2D) / finewidthx
            ry = ry * abs(smooth_scale_2D) / finewidthy
        elif smooth_scale_2D < 1.0:
            rx = smooth_scale_2D * parx.err / finewidthx
            ry = smooth_scale_2D * pary.err / finewidthy
        else:
            rx = smooth_scale_2D * fine_bins_2D / nbin2D
            ry = smooth_scale_2D * fine_bins_2D / nbin2D
        smooth_scale = float(max(rx, ry))
        if smooth_scale < 2:
            log.warning("fine_bins_2D not large enough for optimal density: %s, %s", parx.name, pary.name)
        winw = max(1, int(round(2.5 * smooth_scale)))
        Cinv = np.linalg.inv(np.array([[ry**2, rx * ry * corr], [rx * ry * corr, rx**2]]))
        ix1, ix2 = np.mgrid[-winw: winw + 1, -winw: winw + 1]
        Win = np.exp(-(ix1**2 * Cinv[0, 0] + ix2**2 * Cinv[1, 1] + 2 * Cinv[1, 0] * ix1 * ix2) / 2)
        Win /= np.sum(Win)

        if parx.periodic and pary.periodic:  # mcsamples.py:1874-1882
            cmode = "periodic_both"
        elif parx.periodic:
            cmode = "periodic_x"
        elif pary.periodic:
            cmode = "periodic_y"
        else:
            cmode = "same"
        both_periodic = parx.periodic and pary.periodic
        bins2D = conv2d(histbins, Win, cmode)
        bin2Dlikes = None
        if meanlikes:  # mcsamples.py:1886-1901
            bin2Dlikes = conv2d(finebinlikes, Win, cmode)
            if mult_bias_correction_order:
                lsel = bin2Dlikes > 0
                finebinlikes[lsel] /= bin2Dlikes[lsel]
                likes2 = conv2d(finebinlikes, Win, cmode)
                likes2[lsel] *= bin2Dlikes[lsel]
                bin2Dlikes = likes2
            mxl = 1e-4 * np.max(bins2D)
            bin2Dlikes[bins2D > mxl] /= bins2D[bins2D > mxl]
            bin2Dlikes[bins2D <= mxl] = 0
        prior_mask = None
        bool_mask = None
        if has_prior and boundary_correction_order >= 0 or mult_bias_correction_order or mask_function:
            prior_mask = np.ones((ysize + 2 * winw, xsize + 2 * winw))
            if mask_function:  # mcsamples.py:1909-1919
                mask_function(xbinmin - winw * finewidthx, ybinmin - winw * finewidthy, finewidthx, finewidthy,
                              prior_mask)
                bool_mask = prior_mask[winw:-winw, winw:-winw] < 1e-8
        if has_prior and boundary_correction_order >= 0 and not both_periodic:
            # mcsamples.py:1921-1961 with edge masks :1688-1703 (non-periodic axes only)
            if not parx.periodic:
                if parx.has_limits_bot:
                    prior_mask[:, winw] /= 2
                    prior_mask[:, :winw] = 0
                if parx.has_limits_top:
                    prior_mask[:, -(winw + 1)] /= 2
                    prior_mask[:, -winw:] = 0
            if not pary.periodic:
                if pary.has_limits_bot:
                    prior_mask[winw, :] /= 2
                    prior_mask[:winw:] = 0
                if pary.has_limits_top:
                    prior_mask[-(winw + 1), :] /= 2
                    prior_mask[-winw:, :] = 0
            a00 = conv2d(prior_mask, Win, "valid")
            sel = a00 * bins2D > np.max(bins2D) * 1e-8
            a00 = a00[sel]
            normed = bins2D[sel] / a00
            if boundary_correction_order == 0:
                bins2D[sel] = normed
            elif boundary_correction_order == 1:
                indexes = np.arange(-winw, winw + 1)
                y = np.empty(Win.shape)
                for i in range(Win.shape[0]):
                    y[:, i] = indexes
                winx = Win * indexes
                winy = Win * y
                a10 = conv2d(prior_mask, winx, "valid")[sel]
                a01 = conv2d(prior_mask, winy, "valid")[sel]
                a20 = conv2d(prior_mask, winx * indexes, "valid")[sel]
                a02 = conv2d(prior_mask, winy * y, "valid")[sel]
                a11 = conv2d(prior_mask, winy * indexes, "valid")[sel]
                xP = conv2d(histbins, winx, cmode)[sel]
                yP = conv2d(histbins, winy, cmode)[sel]
                denom = a20 * a01**2 + a10**2 * a02 - a00 * a02 * a20 + a11**2 * a00 - 2 * a01 * a10 * a11
                A = a11**2 - a02 * a20
                Ax = a10 * a02 - a01 * a11
                Ay = a01 * a20 - a10 * a11
                corrected = (bins2D[sel] * A + xP * Ax + yP * Ay) / denom
                bins2D[sel] = normed * np.exp(np.minimum(corrected / normed, 4) - 1)
            else:
                raise OracleSettingError("unknown boundary_correction_order (expected 0 or 1)")
        if mult_bias_correction_order and not both_periodic:
            # mcsamples.py:1963-1976 with :1705-1712 (margins zeroed along non-periodic axes only)
            if not parx.periodic:
                prior_mask[:, :winw] = 0
                prior_mask[:, -winw:] = 0
            if not pary.periodic:
                prior_mask[:winw:] = 0
                prior_mask[-winw:, :] = 0
            a00 = conv2d(prior_mask, Win, "valid")
            for _ in range(mult_bias_correction_order):
                box = histbins.copy()
                sel2 = bins2D > np.max(bins2D) * 1e-8
                box[sel2] /= bins2D[sel2]
                bins2D *= conv2d(box, Win, cmode)
                if mask_function:
                    bins2D[~bool_mask] /= a00[~bool_mask]
                else:
                    bins2D /= a00
        if mask_function:
            bins2D[bool_mask] = 0
        mx = np.max(bins2D)
        if mx == 0:
            raise OracleDensityError("no samples in bin")
        bins2D = bins2D / mx
        x = np.linspace(xbinmin, xbinmax, xsize)
        y = np.linspace(ybinmin, ybinmax, ysize)
        g = Grid2D(x, y, bins2D, ((parx.range_min, parx.range_max), (pary.range_min, pary.range_max)),
                   rx=rx, ry=ry, corr=corr, winw=winw, fine_bins=fine_bins_2D)
        if meanlikes:  # mcsamples.py:2004-2006
            g.likes = bin2Dlikes / np.max(bin2Dlikes)
        g.mask = bool_mask
        g.extra = {k: v for k, v in info.items() if k != "opt"}
        opt = info.get("opt")
        if opt is not None:
            g.extra["t_star"] = opt.t_star
            g.extra["n_brent_evals"] = opt.n_brent_evals
        return g
