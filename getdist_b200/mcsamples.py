"""Host-side mirror of the GetDist ``MCSamples`` hot-path surface, backed by libgdk.so (CUDA, sm_100a).

Mirrors, with the same names, argument meaning and error behaviour (reference paths under
``/root/reference/getdist/``):

    MCSamples.get1DDensity / get1DDensityGridData      mcsamples.py:1500, 1517
    MCSamples.get2DDensity / get2DDensityGridData      mcsamples.py:1730, 1748
    getMeans / getVars / getCov / getCorrelationMatrix chains.py:386, 400, 339, 363
    getGelmanRubin / getGelmanRubinEigenvalues         chains.py:1476, 1446
    confidence / twoTailLimits                         chains.py:814, 782
    updateBaseStatistics / updateSettings              mcsamples.py:552, 472

Everything N-sized or grid-sized runs on the device through the C-ABI (getdist_b200/_abi.py); this module
keeps only the scalar per-parameter logic (``_initParam`` limit tests, grid geometry, 2D branch selection and
2x2 Cholesky algebra), settings handling, caching and error/warning behaviour.  There is no CPU fallback:
the few options of the reference that the device path does not implement (``shade_likes_is_mean_loglikes``,
``range_ND_contour`` with likeStats, ``mask_function`` on periodic parameters) raise ``NotImplementedError``.

``prefetch_triangle`` computes all 1D and 2D densities of a parameter list in batched launches and fills the
caches that the serial ``get1DDensity`` / ``get2DDensity`` calls (as issued by getdist.plots) then hit; a serial
caller that never calls it still gets batches: the first cache miss of ``get1DDensity`` / ``get2DDensityGridData``
triggers a batched launch for the analysis' parameter set (``auto_prefetch``).

Multi-GPU (one process per GPU): ``MCSamples(..., process_group=PeerGroup(...))`` uploads 1/world of the rows per rank
(the peers receive them over NVLink) and ``prefetch_triangle(root=...)`` computes 1/world of the densities per rank;
see getdist_b200/parallel.py and DESIGN.md s6.

Also here: ``getRawNDDensityGridData`` (mcsamples.py:2098-2235) on the histogram kernels and the
MeanVar / Gelman-Rubin / split-quantile parts of ``getConvergeTests`` (mcsamples.py:964-1034) on the per-chain
moment and order-statistics kernels.
"""
import logging
import math
from collections.abc import Mapping

import numpy as np

from . import _abi
from .densities import DensitiesError, Density1D, Density2D, DensityND  # noqa: F401

log = logging.getLogger("getdist_b200")


class WeightedSampleError(Exception):
    pass


class MCSamplesError(WeightedSampleError):
    pass


class SettingError(MCSamplesError):
    pass


class BandwidthError(MCSamplesError):
    pass


class ParamError(MCSamplesError):
    pass


# analysis_defaults.ini of the reference (these override the class attribute defaults)
ANALYSIS_DEFAULTS = dict(
    ignore_rows=0, min_weight_ratio=1e-30, contours=(0.68, 0.95, 0.99), credible_interval_threshold=0.05,
    range_ND_contour=-1, range_confidence=0.001, converge_test_limit=0.95, fine_bins=1024, smooth_scale_1D=-1.0,
    boundary_correction_order=1, mult_bias_correction_order=1, smooth_scale_2D=-1.0, max_corr_2D=0.99,
    fine_bins_2D=256, use_effective_samples_2D=False, max_scatter_points=2000, num_bins=100, num_bins_2D=40,
    num_bins_ND=12,
)

_QFRACS_TAIL = list(np.linspace(0.1, 0.9, 9))


class ParamInfo:
    """Per-parameter state (paramnames.py:69 of the reference; only what the hot path reads/writes)."""

    def __init__(self, name, label=None):
        self.name = name
        self.label = label or ""  # paramnames.py: an unset label stays empty (latexLabel falls back to the name)
        self.isDerived = False
        self.limmin = None
        self.limmax = None
        self.has_limits_bot = False
        self.has_limits_top = False
        self.has_limits = False
        self.periodic = False
        self.N_eff_kde = None
        self.kde_h = None
        self._ranges_ready = False

    def __repr__(self):
        return "ParamInfo(%s)" % self.name


class ParamNames:
    def __init__(self, names, labels=None):
        self.names = [ParamInfo(n, labels[i] if labels else None) for i, n in enumerate(names)]

    def list(self):
        return [p.name for p in self.names]

    def numNonDerived(self):
        return len([p for p in self.names if not p.isDerived])

    def numberOfName(self, name):
        for i, p in enumerate(self.names):
            if p.name == name:
                return i
        return -1

    def parFormat(self):
        """paramnames.py:380-382"""
        return "%-" + str(max(9, max(len(p.name) for p in self.names)) + 1) + "s"


class ParamBounds:
    """Hard prior ranges (parampriors.py:6 of the reference): name -> (lower|None, upper|None)."""

    def __init__(self, ranges=None):
        self.lower, self.upper, self.periodic = {}, {}, set()
        for name, r in (ranges or {}).items():
            self.setRange(name, r)

    def setRange(self, name, r):
        lo, hi = r[0], r[1]
        if lo is not None and not (isinstance(lo, str) and lo == "N"):
            self.lower[name] = float(lo)
        if hi is not None and not (isinstance(hi, str) and hi == "N"):
            self.upper[name] = float(hi)
        if len(r) > 2:
            periodic = r[2]
            if periodic is True or isinstance(periodic, str) and periodic.upper() in ["T", "TRUE", "PERIODIC"]:
                if name not in self.upper or name not in self.lower:
                    raise ValueError("Periodic parameter must have lower and upper bound: %s" % name)
                self.periodic.add(name)

    def getLower(self, name):
        return self.lower.get(name)

    def getUpper(self, name):
        return self.upper.get(name)


class _SpecView:
    """attribute access to one row of the structured gdk_spec2d array"""

    __slots__ = ("_r",)

    def __init__(self, row):
        self._r = row

    def __getattr__(self, k):
        v = self._r[k]
        return v.item() if hasattr(v, "item") and np.ndim(v) == 0 else v


def _overlapped(device_call, host_work):
    """Run `device_call` (a ctypes call into the library: it releases the GIL) in a worker thread while this thread does
    `host_work`; returns both results.  An exception of either side is re-raised here after both have ended."""
    import threading

    box = {}
    running = threading.Event()

    def run():
        running.set()
        try:
            box["r"] = device_call()
        except BaseException as e:  # noqa: BLE001 - handed to the caller's thread
            box["e"] = e

    th = threading.Thread(target=run)
    # The worker must be inside the library call before this thread's Python work takes the interpreter lock: a thread
    # that has to ask for the lock waits a whole switch interval (5 ms by default) each time, and the worker gives the
    # lock up a few times on its way into the call (numpy releases it around array operations) -- measured: 1-11 ms
    # in front of the device work.  So: the worker runs first (event), and hand-overs cost 50 us while both run.
    import sys

    interval = sys.getswitchinterval()
    sys.setswitchinterval(5e-5)
    try:
        th.start()
        running.wait()
        w = host_work()
    finally:
        th.join()
        sys.setswitchinterval(interval)
    if "e" in box:
        raise box["e"]
    return box["r"], w


class _BatchRecord:
    """The library's result record of one density of a batch (bandwidths, status bits, contour levels ...), read like a
    dict: `d._gdk["status"]`.  The columns of the whole batch are shared; nothing is copied per density."""

    __slots__ = ("_cols", "_n", "_conts")

    def __init__(self, cols, n, conts):
        self._cols, self._n, self._conts = cols, n, conts

    def __getitem__(self, key):
        if key == "levels":
            return (self._conts, self._cols["levels"][self._n][: len(self._conts)]) if self._conts else None
        return self._cols[key][self._n]

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __contains__(self, key):
        return key == "levels" or key in self._cols

    def keys(self):
        return [k for k in self._cols] + ["levels"]


class ParamConfidenceData:
    """Handle standing in for chains.ParamConfidenceData: the device resolves order statistics directly, so
    the handle only remembers what it refers to: a stored column (index) with an optional row range, or an
    arbitrary vector with its weights (held in a one-column device context of its own)."""

    def __init__(self, index=None, start=0, end=None, ctx=None):
        self.index = index
        self.start = start
        self.end = end
        self.ctx = ctx


class ChainView:
    """One of the separate chains the samples were combined from (stand-in for the WeightedSamples objects of
    getSeparateChains, chains.py:1505-1527): row range and the per-chain moments of the fused device reduction."""

    def __init__(self, index, start, end, means, cov, norm):
        self.index, self.start, self.end = index, start, end
        self.means, self.fullcov, self.norm = means, cov, norm

    def getMeans(self, pars=None):
        return self.means if pars is None else np.array([self.means[i] for i in pars])

    def getCov(self, nparam=None, pars=None):
        return self.fullcov[np.ix_(pars, pars)] if pars is not None else self.fullcov[:nparam, :nparam]


def _parse_setting(default, v):
    """typed by the default, as IniFile does (inifile.py:216-226): booleans from 'T'/'F' strings, floats stay floats"""
    if isinstance(default, bool):
        if isinstance(v, str):
            return v.strip().lower() in ("t", "true", "1", "yes", "y")
        return bool(v)
    if isinstance(default, int) and not isinstance(v, str) and float(v) != int(v):
        return float(v)  # e.g. a fractional ignore_rows
    return type(default)(v)


def _read_ini(ini):
    """settings from a .ini file name (key = value lines, # comments) or from an object with a `.params` mapping"""
    if hasattr(ini, "params"):
        return dict(ini.params)
    out = {}
    with open(ini) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if "=" in line:
                k, v = line.split("=", 1)
                out[k.strip()] = v.strip()
    return out


class MCSamples:
    def __init__(self, samples=None, weights=None, loglikes=None, names=None, labels=None, ranges=None, sampler=None,
                 settings=None, label=None, device=0, name_tag=None, chain_offsets=None, process_group=None, **kwargs):
        if samples is None:
            raise MCSamplesError("getdist_b200.MCSamples needs in-memory samples (file loading stays in the reference)")
        # chain_offsets: row offsets of already combined chains (chains.py:1497), as the reference object holds them
        self.chain_offsets = None if chain_offsets is None else np.asarray(chain_offsets, dtype=np.int64)
        if isinstance(samples, (list, tuple)) and len(samples) and np.ndim(samples[0]) == 2:
            # chains.py:1488-1503 (makeSingle)
            self.chain_offsets = np.cumsum(np.array([0] + [s.shape[0] for s in samples]))
            if weights is not None:
                weights = np.hstack([np.asarray(w, dtype=np.float64) for w in weights])
            if loglikes is not None:
                loglikes = np.hstack(list(loglikes))
            samples = np.vstack([np.asarray(s, dtype=np.float64) for s in samples])
        samples = np.asarray(samples, dtype=np.float64)
        if samples.ndim == 1:
            samples = samples.reshape(-1, 1)
        self.samples = samples
        self.numrows, self.n = samples.shape
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)
        self.loglikes = None if loglikes is None else np.asarray(loglikes, dtype=np.float64)
        self.mean_loglike = None
        self.shade_likes_is_mean_loglikes = False  # mcsamples.py:233
        self.label = label
        self.name_tag = name_tag
        if names is None:
            names = ["param%d" % (i + 1) for i in range(self.n)]
        if len(names) != self.n:
            raise MCSamplesError("names do not match the number of sample columns")
        self.paramNames = ParamNames(list(names), labels)
        self.index = {p.name: i for i, p in enumerate(self.paramNames.names)}
        self.ranges = ranges if isinstance(ranges, ParamBounds) else ParamBounds(ranges)
        self.sampler = "mcmc" if not isinstance(sampler, str) else sampler.lower()
        if self.sampler not in ("mcmc", "nested", "uncorrelated"):
            self.sampler = "mcmc"
        self.raise_on_bandwidth_errors = False
        self.force_twotail = False  # mcsamples.py:263
        self.no_warning_params = []
        self.no_warning_chi2_params = True
        self.likeStats = None
        self.max_split_tests = 4  # mcsamples.py:262
        self.auto_prefetch = True  # first cache miss of a serial caller triggers a batched launch
        self.auto_prefetch_max_params = 40
        self._seen_2d = []
        for k, v in ANALYSIS_DEFAULTS.items():
            setattr(self, k, v)
        self.contours = np.array(self.contours)
        self._ctx = _abi.Context(device)  # raises if there is no CUDA device / library: no fallback
        # multi-GPU: a getdist_b200.parallel.PeerGroup (one process per GPU); every rank constructs the object with the
        # same arguments, uploads 1/world of the rows and computes 1/world of the densities of prefetch_triangle
        self.process_group = process_group
        self._device_valid = False
        self.density1D = {}
        self._density2D = {}
        self.needs_update = True
        self.updateSettings(settings=settings, doUpdate=False)
        if self.weights is None:
            self.norm = np.float64(self.numrows)
        else:
            self.norm = None
        self.updateBaseStatistics()

    # ------------------------------------------------------------------ pickling / copying
    def __getstate__(self):
        """Picklable / deep-copyable like the reference object (mcsamples.py:125, 316): the device context is not
        part of the state; it is re-created and the samples re-uploaded lazily on first use after loading."""
        d = self.__dict__.copy()
        d["_device"] = self._ctx.device
        d.pop("_ctx", None)
        d["process_group"] = None  # a rendezvous is not part of the state
        d["_device_valid"] = False
        d["_loglikes_valid"] = False
        return d

    def __setstate__(self, d):
        device = d.pop("_device", 0)
        self.__dict__.update(d)
        self._ctx = _abi.Context(device)
        self._device_valid = False
        self.needs_update = True

    def copy(self, label=None, settings=None):
        """mcsamples.py:316-322."""
        import copy as _copy

        new = _copy.deepcopy(self)
        if label is not None:
            new.label = label
        if settings:
            new.updateSettings(settings)
        return new

    # ------------------------------------------------------------------ settings
    def updateSettings(self, settings=None, ini=None, doUpdate=True):
        """mcsamples.py:472-499: `settings` (a dict) override `ini` (a .ini file name or an object with `.params`)."""
        assert settings is None or isinstance(settings, Mapping)
        merged = {}
        if ini is not None:
            merged.update(_read_ini(ini))
        merged.update(settings or {})
        for k, v in merged.items():
            if k == "contours":
                v = np.array(v if not isinstance(v, str) else [float(x) for x in v.split()])
            elif k in ANALYSIS_DEFAULTS and not isinstance(ANALYSIS_DEFAULTS[k], tuple):
                v = _parse_setting(ANALYSIS_DEFAULTS[k], v)
            elif ini is not None and not (settings and k in settings) and not hasattr(self, k):
                continue  # unknown keys of an .ini file are not analysis settings
            setattr(self, k, v)
        # how small the end bin must be relative to the peak for a two-tail limit (mcsamples.py:427-433)
        from scipy.stats import norm as _norm

        self.max_frac_twotail = [float(np.exp(-1.0 * _norm.ppf((1 - c) / 2) ** 2 / 2)) for c in self.contours]
        if doUpdate and self.samples is not None:
            self.updateBaseStatistics()

    # ------------------------------------------------------------------ data residency / statistics
    def _upload(self):
        if not self._device_valid:
            pg = getattr(self, "process_group", None)
            if pg is not None and pg.world > 1 and not pg.probe(self._ctx):
                pg = None  # fallback transport: every rank uploads everything
            self._ctx.set_samples(self.samples, self.weights, self.chain_offsets, **({} if pg is None else {"group": pg}))
            self._device_valid = True
            self._loglikes_valid = False

    def _ensure_loglikes(self):
        """Upload the log-likelihoods once per sample store; the device computes mean_loglike (chains.py:380-381)
        and the mean-likelihood weights weights * exp(mean_loglike - loglikes) (mcsamples.py:1560, 1830)."""
        if self.loglikes is None:
            raise MCSamplesError("meanlikes needs loglikes")
        self._upload()
        if not getattr(self, "_loglikes_valid", False):
            self.mean_loglike = self._ctx.set_loglikes(self.loglikes)
            self._loglikes_valid = True

    def _weightsChanged(self):
        """chains.py:310-323: invalidate everything derived from the samples (and the device copy)."""
        self._device_valid = False
        self.means = None
        self.fullcov = None
        self.correlationMatrix = None
        self.vars = None
        self.sddev = None
        self.needs_update = True

    def setSamples(self, samples, weights=None, loglikes=None):
        """chains.py:262-308; a list of per-chain arrays is combined as in makeSingle (:1488-1503), anything else
        forgets the chain boundaries of the previous samples."""
        self.chain_offsets = None
        if isinstance(samples, (list, tuple)) and len(samples) and np.ndim(samples[0]) == 2:
            self.chain_offsets = np.cumsum(np.array([0] + [s.shape[0] for s in samples]))
            if weights is not None:
                weights = np.hstack([np.asarray(w, dtype=np.float64) for w in weights])
            if loglikes is not None:
                loglikes = np.hstack(list(loglikes))
            samples = np.vstack([np.asarray(s, dtype=np.float64) for s in samples])
        self.samples = np.asarray(samples, dtype=np.float64)
        self.numrows, self.n = self.samples.shape
        self.weights = None if weights is None else np.asarray(weights, dtype=np.float64)
        self.loglikes = None if loglikes is None else np.asarray(loglikes, dtype=np.float64)
        self._weightsChanged()

    def updateBaseStatistics(self):
        """chains.py:1340-1352 + mcsamples.py:552-576: one fused device reduction for means, variances,
        covariance (per chain and total), weight statistics, min/max."""
        self._upload()
        m = self._ctx.moments()
        self._mom = m
        self.means = m["means"]
        self.vars = m["vars"]
        self.sddev = np.sqrt(self.vars)
        self.fullcov = m["cov"]
        self.correlationMatrix = None
        self.norm = m["scalars"][0]
        self.mean_mult = self.norm / self.numrows
        self.max_mult = m["scalars"][2]
        outliers = m["scalars"][3]
        if outliers != 0:
            log.warning("outlier fraction %s ", float(outliers) / self.numrows)
        self._sum_w2 = m["scalars"][1]
        self._xmin, self._xmax = m["xmin"], m["xmax"]
        self.density1D = {}
        self._density2D = {}
        self._seen_2d = []
        self._initLimits()
        for par in self.paramNames.names:
            par.N_eff_kde = None
            par._ranges_ready = False
        self._quantile_cache = {}
        self.needs_update = False
        return self

    def _initLimits(self):
        """mcsamples.py:442-470 (ranges only)."""
        for par in self.paramNames.names:
            par.limmin = self.ranges.getLower(par.name)
            par.limmax = self.ranges.getUpper(par.name)
            par.has_limits_bot = par.limmin is not None
            par.has_limits_top = par.limmax is not None
            par.periodic = par.name in self.ranges.periodic

    def _parAndNumber(self, name):
        """chains.py:1235-1250: name | index | ParamInfo -> (index, ParamInfo); unknown -> (None, None)."""
        if isinstance(name, ParamInfo):
            name = name.name
        if isinstance(name, str):
            ix = self.index.get(name)
            if ix is None:
                return None, None
            return ix, self.paramNames.names[ix]
        ix = int(name)
        if ix < 0 or ix >= self.n:
            return None, None
        return ix, self.paramNames.names[ix]

    def get_norm(self):
        return self.norm

    def getMeans(self, pars=None):
        if self.needs_update:
            self.updateBaseStatistics()
        if pars is None:
            return self.means
        return np.array([self.means[i] for i in pars])

    def getVars(self):
        if self.needs_update:
            self.updateBaseStatistics()
        return self.vars

    def getCov(self, nparam=None, pars=None):
        if self.needs_update:
            self.updateBaseStatistics()
        if pars is not None:
            return self.fullcov[np.ix_(pars, pars)]
        return self.fullcov[:nparam, :nparam]

    def getCorrelationMatrix(self):
        """chains.py:363-371 with covToCorr :155-169."""
        if self.needs_update:
            self.updateBaseStatistics()
        if self.correlationMatrix is None:
            c = self.fullcov.copy()
            for i, di in enumerate(np.sqrt(self.fullcov.diagonal())):
                if di:
                    c[i, :] /= di
                    c[:, i] /= di
            self.correlationMatrix = c
        return self.correlationMatrix

    def getGelmanRubinEigenvalues(self, nparam=None, chainlist=None):
        """chains.py:1446-1474; the per-chain means/covariances come from the fused device reduction, the
        P x P LAPACK work stays on the host."""
        if self.chain_offsets is None:
            raise WeightedSampleError("Samples were not combined from separate chains")
        if self.needs_update:
            self.updateBaseStatistics()
        nparam = nparam or self.paramNames.numNonDerived()
        m = self._mom
        # chainlist: chain indices or the ChainView objects of getSeparateChains() (a subset of the stored chains)
        use = range(m["chain_means"].shape[0]) if chainlist is None else [getattr(c, "index", c) for c in chainlist]
        nch = len(use)
        if nch < 2:
            raise WeightedSampleError("Gelman-Rubin needs at least two chains")
        means = self.means[:nparam]
        meanscov = np.zeros((nparam, nparam))
        meancov = np.zeros((nparam, nparam))
        for c in use:
            diff = m["chain_means"][c, :nparam] - means
            meanscov += np.outer(diff, diff)
            meancov += m["chain_covs"][c, :nparam, :nparam]
        meanscov /= nch - 1
        meancov /= nch
        w, U = np.linalg.eigh(meancov)
        if np.min(w) > 0:
            U /= np.sqrt(w)
            return np.linalg.eigvalsh(np.dot(U.T, meanscov).dot(U))
        return None

    def getGelmanRubin(self, nparam=None, chainlist=None):
        return np.max(self.getGelmanRubinEigenvalues(nparam, chainlist))

    def getSeparateChains(self):
        """chains.py:1505-1527: the separate chains as row ranges with their device-computed moments."""
        if self.chain_offsets is None:
            raise WeightedSampleError("Samples were not combined from separate chains")
        if self.needs_update:
            self.updateBaseStatistics()
        m = self._mom
        return [ChainView(c, int(a), int(b), m["chain_means"][c], m["chain_covs"][c], m["chain_norms"][c])
                for c, (a, b) in enumerate(zip(self.chain_offsets[:-1], self.chain_offsets[1:]))]

    # ------------------------------------------------------------------ order statistics
    def initParamConfidenceData(self, paramVec, start=0, end=None, weights=None):
        """chains.py:793-812.  A stored column (index / name) with the stored weights is addressed in place, with an
        optional row range; an arbitrary vector or other weights are uploaded to a one-column device context."""
        if self.needs_update:
            self.updateBaseStatistics()
        if weights is None and (isinstance(paramVec, (int, np.integer, str, ParamInfo))):
            j, _ = self._parAndNumber(paramVec)
            if j is None:
                raise ParamError("unknown parameter %s" % (paramVec,))
            end = self.numrows if end is None else (end if end >= 0 else self.numrows + end)
            return ParamConfidenceData(j, int(start), int(end))
        if isinstance(paramVec, (int, np.integer, str, ParamInfo)):
            paramVec = self.samples[:, self._parAndNumber(paramVec)[0]]
        vec = np.ascontiguousarray(np.asarray(paramVec, dtype=np.float64)[start:end]).reshape(-1, 1)
        w = self.weights if weights is None else np.asarray(weights, dtype=np.float64)
        w = None if w is None else np.ascontiguousarray(w[start:end])
        ctx = _abi.Context(self._ctx.device)
        ctx.set_samples(vec, w)
        return ParamConfidenceData(ctx=ctx)

    def confidence(self, paramVec, limfrac, upper=False, start=0, end=None, weights=None):
        """chains.py:814-838: sample confidence limits by counting weight in the tails (exact order statistics)."""
        d = paramVec if isinstance(paramVec, ParamConfidenceData) else self.initParamConfidenceData(paramVec, start, end, weights)
        fr = np.atleast_1d(np.asarray(limfrac, dtype=np.float64))
        if upper:
            fr = 1 - fr
        if d.ctx is not None:
            call = lambda f: d.ctx.weighted_quantiles([0], f)[0]  # noqa: E731
        elif d.start == 0 and d.end in (None, self.numrows):
            call = lambda f: self._ctx.weighted_quantiles([d.index], f)[0]  # noqa: E731
        else:
            call = lambda f: self._ctx.weighted_quantiles_range([d.index], f, d.start, d.end)[0]  # noqa: E731
        out = np.concatenate([call(fr[i:i + 16]) for i in range(0, fr.size, 16)])
        return out if np.ndim(limfrac) else out[0]

    def twoTailLimits(self, paramVec, confidence):
        limits = np.array([(1 - confidence) / 2, 1 - (1 - confidence) / 2])
        return self.confidence(paramVec, limits)

    # ------------------------------------------------------------------ N_eff
    def getEffectiveSamplesGaussianKDE(self, paramVec, h=0.2, scale=None, maxoff=None, min_corr=0.05):
        """chains.py:477-574.  Uncorrelated/nested samplers: (sum w)^2 / sum w^2 (:500-501); mcmc: the
        autocorrelation-length / lagged-kernel estimate with every N-sized sum evaluated on the device."""
        if self.sampler in ("nested", "uncorrelated"):
            return self.norm ** 2 / self._sum_w2
        j, _ = self._parAndNumber(paramVec)
        if j is None:
            raise NotImplementedError("device N_eff works on stored columns")
        return self._neff_mcmc_batch([j], [scale], h=h, maxoff=maxoff, min_corr=min_corr)[0]

    def _lag_rounds(self, gens):
        """Drive per-parameter generators in lockstep: each yields a lag-sum request
        (param, mode, k0, nk, mean, inv4s2) and is sent the nk sums; one batched device call per round."""
        out = [None] * len(gens)
        reqs = {}
        for i, g in enumerate(gens):
            try:
                reqs[i] = next(g)
            except StopIteration as e:
                out[i] = e.value
        while reqs:
            order = list(reqs)
            jobs, span = [], []
            for i in order:  # a generator may ask for several blocks of lags at once (a list of requests)
                rq = reqs[i] if isinstance(reqs[i], list) else [reqs[i]]
                span.append((len(jobs), len(rq), isinstance(reqs[i], list)))
                jobs.extend(rq)
            res = self._ctx.lag_sums(jobs)
            nxt = {}
            for i, (a, cnt, many) in zip(order, span):
                try:
                    nxt[i] = gens[i].send(res[a:a + cnt] if many else res[a])
                except StopIteration as e:
                    out[i] = e.value
            reqs = nxt
        return out

    def _corr_length_gen(self, j, min_corr=0.05):
        """getCorrelationLength(d, weight_units=False) (chains.py:448-466) with getAutocorrelation (:423-446):
        corr[k] = sum_i d_i d_{i+k} / (N - k) / var, d = (x - mean) w, for k up to N//10; the reference gets all lags
        from one size-2N FFT (autoConvolve), here lags are produced 16 at a time until the first one that is
        <= min_corr * corr[0]."""
        n = self.numrows
        max_off = n // 10
        mean, var = self.means[j], self.vars[j]
        corr = []
        k0 = 0
        ix = None
        nblk = 1  # blocks of 16 lags per round, growing geometrically: O(log) host round trips for long chains
        while k0 <= max_off and ix is None:
            reqs = []
            kk = k0
            for _ in range(nblk):
                nk = min(16, max_off + 1 - kk)
                if nk <= 0:
                    break
                reqs.append((kk, nk))
                kk += nk
            sums = yield [(j, 0, a, nk, mean, 0.0) for a, nk in reqs]
            for (a, nk), blk in zip(reqs, sums):
                for t in range(nk):
                    k = a + t
                    corr.append(blk[t] / (n - k) / var)
                    if ix is None and not corr[k] > min_corr * corr[0]:
                        ix = k
            k0 = kk
            nblk = min(nblk * 2, 256)
        if ix is None:
            ix = 0  # np.argmin over an all-True array
        return corr[0] + 2 * sum(corr[1:ix])

    def _neff_gen(self, j, scale, h, maxoff, min_corr):
        """getEffectiveSamplesGaussianKDE, mcmc branch (chains.py:502-574)."""
        n = self.numrows
        kernel_std = (scale or self.sddev[j]) * h
        inv4 = 1.0 / (4 * kernel_std ** 2)
        if maxoff is None:
            clen = yield from self._corr_length_gen(j)
            maxoff = int(clen * 1.5) + 4
        maxoff = min(maxoff, n // 10)
        uncorr_len = n // 2
        sums = yield (j, 1, uncorr_len, 5, 0.0, inv4)
        nav = sum(n - k for k in range(uncorr_len, uncorr_len + 5))
        uncorr_term = float(np.sum(sums)) / nav
        corr0 = self._sum_w2
        nn = float(n)

        def corr_k(k):
            s = yield (j, 1, k, 1, 0.0, inv4)
            return s[0] - (nn - k) * uncorr_term

        threshold = min_corr * corr0
        c1 = yield from corr_k(1)
        if c1 < threshold:
            N = corr0
        else:
            c2 = yield from corr_k(2)
            if c2 > threshold:
                max_k = maxoff
                while max_k > 10:
                    test_val = yield from corr_k(max_k // 3)
                    if test_val >= threshold:
                        break
                    max_k //= 3
                step_size = 1 if max_k < 20 else max_k // 10
                cum_sum = c1 + c2
                for k in range(3, maxoff + 1, step_size):
                    test_val = yield from corr_k(k)
                    if test_val < threshold:
                        break
                    if k > 3:
                        cum_sum += test_val * step_size
                    else:
                        cum_sum += (test_val * step_size) / 2
                N = corr0 + 2 * cum_sum
            else:
                N = corr0 + 2 * c1
        return self.norm ** 2 / N

    def _neff_mcmc_batch(self, indices, scales, h=0.2, maxoff=None, min_corr=0.05):
        if self.needs_update:
            self.updateBaseStatistics()
        gens = [self._neff_gen(j, sc, h, maxoff, min_corr) for j, sc in zip(indices, scales)]
        return self._lag_rounds(gens)

    def getCorrelationLength(self, j, weight_units=True, min_corr=0.05, corr=None):
        """chains.py:448-466.  Weight units scale every lag of the autocorrelation by numrows / sum(w)
        (getAutocorrelation, :443-444), so the threshold crossing is the same and the length scales likewise."""
        if corr is not None:  # an autocorrelation array handed in: plain host arithmetic on it
            corr = np.asarray(corr)
            ix = np.argmin(corr > min_corr * corr[0])
            return corr[0] + 2 * np.sum(corr[1:ix])
        if self.needs_update:
            self.updateBaseStatistics()
        j, _ = self._parAndNumber(j)
        n = self._lag_rounds([self._corr_length_gen(j, min_corr)])[0]
        return n * self.numrows / self.norm if weight_units else n

    def _get1DNeff(self, par, param):
        if par.N_eff_kde is None:
            par.N_eff_kde = self.getEffectiveSamplesGaussianKDE(param, scale=par.sigma_range)
        return par.N_eff_kde

    def _ensure_neff(self, indices):
        """N_eff for every listed parameter that does not have one yet; mcmc sampler: batched rounds."""
        todo = [j for j in dict.fromkeys(indices) if self.paramNames.names[j].N_eff_kde is None]
        if not todo:
            return
        if self.sampler in ("nested", "uncorrelated"):
            for j in todo:
                self.paramNames.names[j].N_eff_kde = self.norm ** 2 / self._sum_w2
        else:
            vals = self._neff_mcmc_batch(todo, [self.paramNames.names[j].sigma_range for j in todo])
            for j, v in zip(todo, vals):
                self.paramNames.names[j].N_eff_kde = v

    # ------------------------------------------------------------------ parameter ranges
    def _range_fracs(self):
        """the 11 probability fractions _initParam asks confidence() for (mcsamples.py:1440-1443)"""
        return np.array([self.range_confidence, 1 - self.range_confidence] + _QFRACS_TAIL)

    def _ensure_param_ranges(self, indices):
        """_initParam (mcsamples.py:1427-1484) for a set of parameters: one batched exact-quantile call for
        those not done yet, then the scalar range / limit logic per parameter."""
        todo = [j for j in dict.fromkeys(indices) if not self.paramNames.names[j]._ranges_ready]
        if not todo:
            return
        fr = self._range_fracs()
        self._finish_params(todo, self._ctx.weighted_quantiles(todo, fr))

    def _finish_param(self, par, j, confids):
        self._finish_params([j], [confids])

    def _finish_params(self, js, table):
        """The scalar range / limit logic of _initParam (mcsamples.py:1438-1484) for a set of parameters at once:
        row k of `table` holds the 11 order statistics of parameter js[k] (_range_fracs).  Written as array
        expressions over the parameters -- the same IEEE operations per element as the per-parameter statements of
        the reference -- because a 64-parameter triangle pays for 64 x a dozen tiny numpy calls otherwise."""
        if self.range_ND_contour >= 0 and self.likeStats:
            raise NotImplementedError("range_ND_contour needs likeStats (mcsamples.py:1455-1459): use the reference")
        js = np.asarray(js, dtype=np.int64)
        pars = [self.paramNames.names[j] for j in js]
        q = np.asarray(table, dtype=np.float64).reshape(len(pars), -1)
        err, mean, pmin, pmax = self.sddev[js], self.means[js], self._xmin[js], self._xmax[js]
        rmin, rmax = q[:, 0], q[:, 1]
        conf = np.empty((len(pars), q.shape[1]))  # [param_min, the tail quantiles, param_max]
        conf[:, 0], conf[:, 1:-1], conf[:, -1] = pmin, q[:, 2:], pmax
        diffs = conf[:, 4:] - conf[:, :-4]
        scale = diffs.min(axis=1) / 1.049
        regular = np.all(diffs > (err * 1.049)[:, None], axis=1) & np.all(diffs < (scale * 1.5)[:, None], axis=1)
        sigma = np.where(regular, scale, np.minimum(err, scale))
        smooth = sigma * 0.4
        lo = np.array([np.nan if p.limmin is None else p.limmin for p in pars], dtype=np.float64)
        hi = np.array([np.nan if p.limmax is None else p.limmax for p in pars], dtype=np.float64)
        with np.errstate(invalid="ignore"):
            # a hard limit far outside the samples is dropped (mcsamples.py:1465-1476), otherwise the range snaps to it
            bot = ~np.isnan(lo) & ~((rmin - lo > 2 * smooth) & (pmin - lo > smooth))
            top = ~np.isnan(hi) & ~((hi - rmax > 2 * smooth) & (hi - pmax > smooth))
        rmin = np.where(bot, lo, rmin - smooth * 2)
        rmax = np.where(top, hi, rmax + smooth * 2)
        for k, par in enumerate(pars):
            par.err, par.mean, par.param_min, par.param_max = err[k], mean[k], pmin[k], pmax[k]
            par.sigma_range = sigma[k]
            par.range_min, par.range_max = rmin[k], rmax[k]
            par.has_limits_bot, par.has_limits_top = bool(bot[k]), bool(top[k])
            par.has_limits = par.has_limits_bot or par.has_limits_top
            par._ranges_ready = True

    def _initParamRanges(self, j, paramConfid=None):
        j, par = self._parAndNumber(j)
        self._ensure_param_ranges([j])
        return par

    @staticmethod
    def _bin_geometry(par, num_fine_bins, borderfrac=0.1):
        """mcsamples.py:1486-1496."""
        border = (par.range_max - par.range_min) * borderfrac
        binmin = min(par.param_min, par.range_min)
        if not par.has_limits_bot:
            binmin -= border
        binmax = max(par.param_max, par.range_max)
        if not par.has_limits_top:
            binmax += border
        return binmin, binmax

    # ------------------------------------------------------------------ 1D densities
    def get1DDensity(self, name, **kwargs):
        """mcsamples.py:1500-1514."""
        if self.needs_update:
            self.updateBaseStatistics()
        if not kwargs:
            j, par = self._parAndNumber(name)
            if par is not None:
                density = self.density1D.get(par.name)
                if density is None and self.auto_prefetch and self.n <= 4 * self.auto_prefetch_max_params:
                    # first miss of a serial caller: all 1D densities in one batched launch
                    self._densities_1d([k for k in range(self.n) if self.paramNames.names[k].name not in self.density1D])
                    density = self.density1D.get(par.name)
                if density is not None:
                    return density
        return self.get1DDensityGridData(name, **kwargs)

    def get1DDensityGridData(self, j, paramConfid=None, meanlikes=False, **kwargs):
        """mcsamples.py:1517-1686."""
        if self.needs_update:
            self.updateBaseStatistics()
        j = self._parAndNumber(j)[0]
        if j is None:
            return None
        return self._densities_1d([j], meanlikes=meanlikes, **kwargs)[0]

    def _spec_1d(self, j, kwargs):
        par = self.paramNames.names[j]
        num_bins = kwargs.get("num_bins", self.num_bins)
        smooth_scale_1D = kwargs.get("smooth_scale_1D", self.smooth_scale_1D)
        bco = kwargs.get("boundary_correction_order", self.boundary_correction_order)
        mbc = kwargs.get("mult_bias_correction_order", self.mult_bias_correction_order)
        fine_bins = int(kwargs.get("fine_bins", self.fine_bins))
        paramrange = par.range_max - par.range_min
        if paramrange <= 0:
            raise MCSamplesError("Parameter range is <= 0: " + par.name)
        if par.has_limits and not par.periodic and bco > 2:
            raise SettingError("Unknown boundary_correction_order (expected 0, 1, 2)")
        width = paramrange / (num_bins - 1)
        binmin, binmax = self._bin_geometry(par, fine_bins)
        neff = self._get1DNeff(par, j) if smooth_scale_1D <= 0 else 1.0
        return _abi.Spec1D(j, fine_bins, binmin, binmax, par.range_min, par.range_max, par.param_min, par.param_max,
                           par.sigma_range, par.err, neff, float(smooth_scale_1D), width, int(bco), int(mbc),
                           int(par.has_limits_bot), int(par.has_limits_top), int(par.periodic), 0)

    def _densities_1d(self, indices, meanlikes=False, _out=None, _device_ptr=None, **kwargs):
        if meanlikes:
            if self.shade_likes_is_mean_loglikes:
                raise NotImplementedError("shade_likes_is_mean_loglikes (signed weights * loglikes) is not on the device path")
            self._ensure_loglikes()
        self._ensure_param_ranges(indices)
        if kwargs.get("smooth_scale_1D", self.smooth_scale_1D) <= 0:
            self._ensure_neff(indices)
        specs = [self._spec_1d(j, kwargs) for j in indices]
        L = None
        if meanlikes:
            P, L, res = self._ctx.density1d_batch(specs, likes=True)
        else:
            P, res = self._ctx.density1d_batch(specs, out=_out, device_ptr=_device_ptr)
        if _device_ptr is not None:
            return specs, res  # grids stay on the device (row i at device_ptr + i * max(fine_bins))
        return self._finish_1d(indices, specs, P, res, L, cache=not kwargs)

    def _finish_1d(self, indices, specs, P, res, L=None, cache=True):
        """Density1D objects (+ the reference's warnings / errors and ParamInfo side effects) from the rows of a 1D batch"""
        out = []
        for k, (j, sp, row, r) in enumerate(zip(indices, specs, P, res)):
            par = self.paramNames.names[j]
            if sp.smooth_scale_1D <= 0:
                if r.status & _abi.ST_BW_FALLBACK:
                    # mcsamples.py:1258-1268
                    if par.name not in self.no_warning_params and (
                            not self.no_warning_chi2_params or "chi2_" not in par.name and "minuslog" not in par.name):
                        h = None if (r.status & _abi.ST_BW_FAILED_NONE) else r.h_raw
                        msg = "auto bandwidth for %s very small or failed (h=%s,N_eff=%s). Using fallback (h=%s)" % (
                            par.name, h, sp.neff, r.kde_h)
                        if self.raise_on_bandwidth_errors:
                            raise BandwidthError(msg)
                        log.warning(msg)
                par.kde_h = r.kde_h
            if r.status & _abi.ST_SMALL_SMOOTH:
                log.warning("fine_bins not large enough to well sample smoothing scale - " + par.name)
            if r.status & _abi.ST_ZERO_MAX:
                raise DensitiesError("no samples in bin")
            x = np.linspace(sp.binmin, sp.binmax, sp.fine_bins)
            d = Density1D(x, P=row[: sp.fine_bins], view_ranges=[par.range_min, par.range_max])
            d.likes = None if L is None else L[k][: sp.fine_bins]  # mcsamples.py:1672-1684
            d._gdk = dict(kde_h=r.kde_h, smooth_1D=r.smooth_1D, winw=r.winw, status=r.status, n_feval=r.n_feval)
            if cache:
                self.density1D[par.name] = d
            out.append(d)
        return out

    # ------------------------------------------------------------------ 2D densities
    def get2DDensity(self, x, y, normalized=False, **kwargs):
        """mcsamples.py:1730-1745."""
        if self.needs_update:
            self.updateBaseStatistics()
        density = self.get2DDensityGridData(x, y, get_density=True, **kwargs)
        if normalized:
            density.normalize(in_place=True)
        return density

    def get2DDensityGridData(self, j, j2, num_plot_contours=None, get_density=False, meanlikes=False,
                             mask_function=None, **kwargs):
        """mcsamples.py:1748-2010."""
        if self.needs_update:
            self.updateBaseStatistics()
        j, parx = self._parAndNumber(j)
        j2, pary = self._parAndNumber(j2)
        if j is None or j2 is None:
            return None
        ncontours = len(self.contours)
        if num_plot_contours:
            ncontours = min(num_plot_contours, ncontours)
        want = [] if get_density else list(self.contours[:ncontours])
        density = None
        if meanlikes or mask_function is not None:
            # mcsamples.py:1829-1831, 1886-1901, 2004-2006 (likes are only attached when get_density=False);
            # mask_function: mcsamples.py:1909-1919, 1973-1979
            if mask_function is not None and (parx.periodic or pary.periodic):
                raise NotImplementedError("mask_function with periodic parameters is not on the device path")
            density = self._densities_2d([(j, j2)], _contours=want, _likes=bool(meanlikes), _mask_function=mask_function, **kwargs)[0]
        elif not kwargs:
            if (j, j2) not in self._density2D and self.auto_prefetch:
                self._auto_prefetch_2d(j, j2)
            density = self._cached_2d(j, j2)
        if density is None:
            density = self._densities_2d([(j, j2)], _contours=want, **kwargs)[0]
            if not kwargs and not meanlikes and mask_function is None:
                density = self._cached_2d(j, j2)  # the cache keeps a private copy; callers get a fresh object
        if get_density:
            return density
        dev = density._gdk.get("levels")
        if (dev is not None and len(dev[0]) >= len(want) and list(dev[0][:len(want)]) == want
                and not density._gdk["status"] & _abi.ST_CONTOUR_RANGE):
            density.contours = np.array(dev[1][:len(want)])
        else:
            # more than 4 contours, other fractions than the cached ones, or one of the device levels fell outside the
            # plotted range (one status bit per density): the host evaluates exactly the requested ones and raises
            # DensitiesError only if one of THOSE is out of range (densities.py:50-51)
            density.contours = density.getContourLevels(want)
        density.likes = getattr(density, "_likes2d", None) if meanlikes else None
        return density

    def _cached_2d(self, j, j2):
        """a fresh Density2D (own copy of the grid) from the cache entry of the pair, or None: callers normalise in
        place (get2DDensity(normalized=True)), the reference returns a new object per call"""
        c = self._density2D.get((j, j2))
        if c is None:
            return None
        d = Density2D(c.x, c.y, c.P.copy(), view_ranges=c.view_ranges)
        d._gdk = c._gdk
        return d

    def _auto_prefetch_2d(self, j, j2):
        """First 2D cache miss of a serial caller (getdist.plots.triangle_plot asks pair by pair, plots.py:2613):
        compute a batch instead of one pair.  Few parameters: the whole triangle at once.  Many parameters: every
        uncached pair among the parameters seen so far -- for the triangle order that is one plot row per launch."""
        if self.n <= self.auto_prefetch_max_params:
            idx = list(range(self.n))
            pairs = [(a, b) for a in idx for b in idx if a != b and ((a, b) not in self._density2D) and a < b]
            if (j, j2) not in pairs:
                pairs.append((j, j2))
        else:
            for q in (j, j2):
                if q not in self._seen_2d:
                    self._seen_2d.append(q)
            # same orientation as the request: x = the earlier seen parameter unless asked otherwise
            pairs = [(j, j2)]
            for a in self._seen_2d:
                if a in (j, j2):
                    continue
                for pr in ((a, j2), (a, j)):
                    if pr not in self._density2D and (pr[1], pr[0]) not in self._density2D and pr not in pairs and pr[0] != pr[1]:
                        pairs.append(pr)
        self._densities_2d(pairs)

    def _spec_2d(self, j, j2, kwargs):
        parx, pary = self.paramNames.names[j], self.paramNames.names[j2]
        base_fine_bins_2D = int(kwargs.get("fine_bins_2D", self.fine_bins_2D))
        bco = kwargs.get("boundary_correction_order", self.boundary_correction_order)
        mbc = kwargs.get("mult_bias_correction_order", self.mult_bias_correction_order)
        smooth_scale_2D = float(kwargs.get("smooth_scale_2D", self.smooth_scale_2D))
        has_prior = parx.has_limits or pary.has_limits
        corr = self.getCorrelationMatrix()[j2][j]
        actual_corr = corr
        if abs(abs(corr) - 1.0) <= 1e-8:
            log.warning("Parameters are 100%% correlated: %s, %s", parx.name, pary.name)
            corr = np.sign(corr) * self.max_corr_2D
        if abs(self.max_corr_2D) > 1:
            raise SettingError("max_corr_2D cannot be >=1")
        if abs(corr) < 0.1:
            corr = 0.0
        if has_prior and bco > 1 and not (parx.periodic and pary.periodic):
            raise SettingError("unknown boundary_correction_order (expected 0 or 1)")
        angle_scale = max(0.2, np.sqrt(1 - min(self.max_corr_2D, abs(corr)) ** 2))
        nbin2D = int(round(self.num_bins_2D / angle_scale))
        fine_bins_2D = base_fine_bins_2D
        if corr:
            scaled = 192 * int(3 / angle_scale) // 3
            if base_fine_bins_2D < scaled and int(1 / angle_scale) > 1:
                fine_bins_2D = scaled
        xbinmin, xbinmax = self._bin_geometry(parx, fine_bins_2D)
        ybinmin, ybinmax = self._bin_geometry(pary, fine_bins_2D)
        finewidthx = (xbinmax - xbinmin) / (fine_bins_2D - 1)
        finewidthy = (ybinmax - ybinmin) / (fine_bins_2D - 1)
        sp = _abi.Spec2D()
        sp.px, sp.py = j, j2
        sp.fine_bins, sp.base_fine_bins = fine_bins_2D, base_fine_bins_2D
        sp.xbinmin, sp.xbinmax, sp.ybinmin, sp.ybinmax = xbinmin, xbinmax, ybinmin, ybinmax
        sp.x_sigma_range, sp.y_sigma_range = parx.sigma_range, pary.sigma_range
        sp.x_err, sp.y_err = parx.err, pary.err
        sp.corr = actual_corr
        sp.kernel_corr = corr
        sp.max_corr_2D = self.max_corr_2D
        sp.smooth_scale_2D = smooth_scale_2D
        sp.boundary_correction_order = int(bco)
        sp.mult_bias_correction_order = int(mbc)
        sp.x_has_bot, sp.x_has_top = int(parx.has_limits_bot), int(parx.has_limits_top)
        sp.y_has_bot, sp.y_has_top = int(pary.has_limits_bot), int(pary.has_limits_top)
        sp.x_periodic, sp.y_periodic = int(parx.periodic), int(pary.periodic)
        sp.neff = 1.0
        if smooth_scale_2D < 0:
            # branch selection of getAutoBandwidth2D, mcsamples.py:1325-1409
            # use_effective_samples_2D is inert on this path in the reference (getAutoBandwidth2D's use_2D_Neff=False
            # default shadows the setting, mcsamples.py:1297, 1327-1331; verified by running it): always the 1D estimate
            sp.neff = min(self._get1DNeff(parx, j), self._get1DNeff(pary, j2))
            do_correlated = not parx.has_limits or not pary.has_limits
            min_corr = 0.2
            if min_corr < abs(actual_corr) <= self.max_corr_2D and do_correlated:
                sp.bw_mode = _abi.BW2D_SHEAR
                i, k = j, j2
                imax, imin = None, None
                if parx.has_limits_bot:
                    imin = parx.range_min
                if parx.has_limits_top:
                    imax = parx.range_max
                swapped = 0
                if pary.has_limits:
                    i, k = k, i
                    swapped = 1
                    if pary.has_limits_bot:
                        imin = pary.range_min
                    if pary.has_limits_top:
                        imax = pary.range_max
                cov = self.getCov(pars=[i, k])
                S = np.linalg.cholesky(cov)
                ichol = np.linalg.inv(S)
                S = S * ichol[0, 0]
                r = ichol[1, :] / ichol[0, 0]
                sp.shear_i, sp.shear_j, sp.shear_swapped = i, k, swapped
                sp.r0, sp.r1 = r[0], r[1]
                sp.S00, sp.S10, sp.S11 = S[0, 0], S[1, 0], S[1, 1]
                # kde.bin_samples range of p1 (kde_bandwidth.py:77-84)
                mn, mx = self._xmin[i], self._xmax[i]
                delta = mx - mn
                sp.p1_min = imin if imin is not None else mn - delta * 0.1
                sp.p1_max = imax if imax is not None else mx + delta * 0.1
            elif abs(actual_corr) > self.max_corr_2D or not do_correlated and actual_corr > 0.8:
                sp.bw_mode = _abi.BW2D_RULE
            else:
                sp.bw_mode = _abi.BW2D_PLAIN
        else:
            sp.bw_mode = _abi.BW2D_FIXED
            if smooth_scale_2D < 1.0:
                sp.rx_fixed = smooth_scale_2D * parx.err / finewidthx
                sp.ry_fixed = smooth_scale_2D * pary.err / finewidthy
            else:
                sp.rx_fixed = smooth_scale_2D * fine_bins_2D / nbin2D
                sp.ry_fixed = smooth_scale_2D * fine_bins_2D / nbin2D
        return sp

    def _specs_2d_batch(self, pairs, kwargs):
        """Vectorised equivalent of [_spec_2d(j, j2, kwargs) for (j, j2) in pairs]: the same IEEE operations on
        numpy arrays (the 2x2 Cholesky / inverse of the shear branch as stacked LAPACK calls), written straight
        into a structured array with the layout of gdk_spec2d.  Keeps the host out of the critical path of a
        2016-pair triangle."""
        n = len(pairs)
        names = self.paramNames.names
        jx = np.fromiter((p[0] for p in pairs), dtype=np.int64, count=n)
        jy = np.fromiter((p[1] for p in pairs), dtype=np.int64, count=n)

        used = np.zeros(len(names), dtype=bool)
        used[jx] = True
        used[jy] = True

        def col(a, dt=np.float64):  # parameters outside this batch may not have their ranges yet
            return np.array([getattr(par, a) if u else 0 for par, u in zip(names, used)], dtype=dt)

        rmin, rmax, pmin, pmax = col("range_min"), col("range_max"), col("param_min"), col("param_max")
        sig, err = col("sigma_range"), col("err")
        hb, ht, per = col("has_limits_bot", bool), col("has_limits_top", bool), col("periodic", bool)
        hl = hb | ht
        base = int(kwargs.get("fine_bins_2D", self.fine_bins_2D))
        bco = int(kwargs.get("boundary_correction_order", self.boundary_correction_order))
        mbc = int(kwargs.get("mult_bias_correction_order", self.mult_bias_correction_order))
        smooth = float(kwargs.get("smooth_scale_2D", self.smooth_scale_2D))
        if abs(self.max_corr_2D) > 1:
            raise SettingError("max_corr_2D cannot be >=1")
        actual = self.getCorrelationMatrix()[jy, jx]
        corr = actual.copy()
        one = np.abs(np.abs(corr) - 1.0) <= 1e-8
        for k in np.nonzero(one)[0]:
            log.warning("Parameters are 100%% correlated: %s, %s", names[jx[k]].name, names[jy[k]].name)
        corr[one] = np.sign(corr[one]) * self.max_corr_2D
        corr[np.abs(corr) < 0.1] = 0.0
        has_prior = hl[jx] | hl[jy]
        if bco > 1 and np.any(has_prior & ~(per[jx] & per[jy])):
            raise SettingError("unknown boundary_correction_order (expected 0 or 1)")
        angle = np.maximum(0.2, np.sqrt(1 - np.minimum(self.max_corr_2D, np.abs(corr)) ** 2))
        nbin2D = np.rint(self.num_bins_2D / angle).astype(np.int64)
        scaled = 192 * (3 / angle).astype(np.int64) // 3
        fine = np.where((corr != 0) & (base < scaled) & ((1 / angle).astype(np.int64) > 1), scaled, base).astype(np.int64)

        def geom(j):
            border = (rmax[j] - rmin[j]) * 0.1
            lo = np.minimum(pmin[j], rmin[j])
            lo = np.where(hb[j], lo, lo - border)
            hi = np.maximum(pmax[j], rmax[j])
            hi = np.where(ht[j], hi, hi + border)
            return lo, hi

        xlo, xhi = geom(jx)
        ylo, yhi = geom(jy)
        fwx = (xhi - xlo) / (fine - 1)
        fwy = (yhi - ylo) / (fine - 1)
        sp = np.zeros(n, dtype=_abi.SPEC2D_DTYPE)
        sp["px"], sp["py"] = jx, jy
        sp["fine_bins"], sp["base_fine_bins"] = fine, base
        sp["xbinmin"], sp["xbinmax"], sp["ybinmin"], sp["ybinmax"] = xlo, xhi, ylo, yhi
        sp["x_sigma_range"], sp["y_sigma_range"] = sig[jx], sig[jy]
        sp["x_err"], sp["y_err"] = err[jx], err[jy]
        sp["corr"], sp["kernel_corr"] = actual, corr
        sp["max_corr_2D"] = self.max_corr_2D
        sp["smooth_scale_2D"] = smooth
        sp["boundary_correction_order"], sp["mult_bias_correction_order"] = bco, mbc
        sp["x_has_bot"], sp["x_has_top"], sp["y_has_bot"], sp["y_has_top"] = hb[jx], ht[jx], hb[jy], ht[jy]
        sp["x_periodic"], sp["y_periodic"] = per[jx], per[jy]
        sp["neff"] = 1.0
        if smooth < 0:
            neff = np.array([par.N_eff_kde if u else np.nan for par, u in zip(names, used)], dtype=np.float64)
            sp["neff"] = np.minimum(neff[jx], neff[jy])
            do_corr = ~hl[jx] | ~hl[jy]
            shear = (0.2 < np.abs(actual)) & (np.abs(actual) <= self.max_corr_2D) & do_corr
            rule = ~shear & ((np.abs(actual) > self.max_corr_2D) | (~do_corr & (actual > 0.8)))
            mode = np.where(shear, _abi.BW2D_SHEAR, np.where(rule, _abi.BW2D_RULE, _abi.BW2D_PLAIN))
            sp["bw_mode"] = mode
            ks = np.nonzero(shear)[0]
            if ks.size:
                swapped = hl[jy[ks]]  # pary.has_limits: (i, j) = (py, px)
                si = np.where(swapped, jy[ks], jx[ks])
                sk = np.where(swapped, jx[ks], jy[ks])
                # imin / imax: limits of x, overridden by those of y when y has limits (mcsamples.py:1352-1362)
                imin = np.where(hb[jx[ks]], rmin[jx[ks]], np.nan)
                imax = np.where(ht[jx[ks]], rmax[jx[ks]], np.nan)
                imin = np.where(swapped & hb[jy[ks]], rmin[jy[ks]], imin)
                imax = np.where(swapped & ht[jy[ks]], rmax[jy[ks]], imax)
                cov = self.fullcov
                covs = np.empty((ks.size, 2, 2))
                covs[:, 0, 0], covs[:, 0, 1] = cov[si, si], cov[si, sk]
                covs[:, 1, 0], covs[:, 1, 1] = cov[sk, si], cov[sk, sk]
                S = np.linalg.cholesky(covs)
                ichol = np.linalg.inv(S)
                S = S * ichol[:, 0:1, 0:1]
                r = ichol[:, 1, :] / ichol[:, 0, 0][:, None]
                sp["shear_i"][ks], sp["shear_j"][ks], sp["shear_swapped"][ks] = si, sk, swapped
                sp["r0"][ks], sp["r1"][ks] = r[:, 0], r[:, 1]
                sp["S00"][ks], sp["S10"][ks], sp["S11"][ks] = S[:, 0, 0], S[:, 1, 0], S[:, 1, 1]
                mn, mx = self._xmin[si], self._xmax[si]
                delta = mx - mn
                sp["p1_min"][ks] = np.where(np.isnan(imin), mn - delta * 0.1, imin)
                sp["p1_max"][ks] = np.where(np.isnan(imax), mx + delta * 0.1, imax)
        else:
            sp["bw_mode"] = _abi.BW2D_FIXED
            if smooth < 1.0:
                sp["rx_fixed"] = smooth * err[jx] / fwx
                sp["ry_fixed"] = smooth * err[jy] / fwy
            else:
                sp["rx_fixed"] = smooth * fine / nbin2D
                sp["ry_fixed"] = smooth * fine / nbin2D
        return sp

    def _fine_bins_2d_all(self, pairs, jx=None, jy=None):
        """fine_bins of every pair (mcsamples.py:1811-1818: scaled up for strongly correlated pairs): from the correlation
        matrix and the settings alone (jx / jy: the pairs' columns as index arrays, where the caller keeps them)"""
        if jx is None:
            jx = np.array([p[0] for p in pairs], dtype=np.int64)
            jy = np.array([p[1] for p in pairs], dtype=np.int64)
        base = int(self.fine_bins_2D)
        corr = self.getCorrelationMatrix()[jy, jx].copy()
        one = np.abs(np.abs(corr) - 1.0) <= 1e-8
        corr[one] = np.sign(corr[one]) * self.max_corr_2D
        corr[np.abs(corr) < 0.1] = 0.0
        angle = np.maximum(0.2, np.sqrt(1 - np.minimum(self.max_corr_2D, np.abs(corr)) ** 2))
        scaled = 192 * (3 / angle).astype(np.int64) // 3
        return np.where((corr != 0) & (base < scaled) & ((1 / angle).astype(np.int64) > 1), scaled, base).astype(np.int64)

    def _densities_2d(self, pairs, _out=None, _device_ptr=None, _contours=None, _likes=False, _anchor_hints=None,
                      _mask_function=None, **kwargs):
        if _likes:
            self._ensure_loglikes()
        self._ensure_param_ranges([p for pr in pairs for p in pr])
        if float(kwargs.get("smooth_scale_2D", self.smooth_scale_2D)) < 0:
            self._ensure_neff([p for pr in pairs for p in pr])
        specs = self._specs_2d_batch(pairs, kwargs)
        if _contours is None:
            _contours = list(self.contours[:4])  # batched prefetch: the analysis-settings contours
        conts = [float(c) for c in _contours[:4]] if len(_contours) <= 4 else []
        specs["n_contours"] = len(conts)
        if _anchor_hints is not None:
            specs["anchor_hint"] = np.asarray(_anchor_hints, dtype=np.int32)
        for k, c in enumerate(conts):
            specs["contours"][:, k] = c
        lbuf = None
        masks = None
        if _mask_function is not None:
            # the prior mask needs the kernel half-width (mcsamples.py:1863, 1909-1916): bandwidth stage first, then the
            # full pipeline with that bandwidth fixed and the (G + 2w)^2 masks the user function filled in
            res1 = self._ctx.bandwidth2d_batch(specs)
            masks, mask_w = [], []
            for spr, r in zip(specs, res1):
                G, w = int(spr["fine_bins"]), int(r.winw)
                fwx = (float(spr["xbinmax"]) - float(spr["xbinmin"])) / (G - 1)
                fwy = (float(spr["ybinmax"]) - float(spr["ybinmin"])) / (G - 1)
                pm = np.ones((G + 2 * w, G + 2 * w))
                _mask_function(float(spr["xbinmin"]) - w * fwx, float(spr["ybinmin"]) - w * fwy, fwx, fwy, pm)
                masks.append(pm)
                mask_w.append(w)
            specs2 = specs.copy()
            for k, r in enumerate(res1):
                if int(specs2["bw_mode"][k]) != 0:
                    specs2["bw_mode"][k] = 0  # GDK_BW2D_FIXED: the widths (in bins) and correlation found above
                    specs2["rx_fixed"][k], specs2["ry_fixed"][k], specs2["kernel_corr"][k] = r.rx, r.ry, r.c
            buf, lbuf, offsets, res = self._ctx.density2d_masked_batch(specs2, masks, mask_w, likes=_likes)
            for r2, r in zip(res, res1):  # bandwidth diagnostics and warnings come from the first stage
                r2.status |= r.status
                r2.hx, r2.hy, r2.t_star, r2.n_brent = r.hx, r.hy, r.t_star, r.n_brent
        elif _likes:
            buf, lbuf, offsets, res = self._ctx.density2d_batch(specs, likes=True)
        elif _device_ptr is not None:
            buf, offsets, res = self._ctx.density2d_batch(specs, device_ptr=_device_ptr)
            return specs, offsets, res  # grids stay on the device (density i at device_ptr + offsets[i])
        elif len(pairs) >= 64 and hasattr(_abi, "result_buffer"):
            # large batch: the host wraps the grids (views of the result buffer) while the library call is in flight
            fb = specs["fine_bins"].astype(np.int64)
            offsets = np.zeros(len(pairs), dtype=np.int64)
            offsets[1:] = np.cumsum(fb * fb)[:-1]
            buf = _out if _out is not None else _abi.result_buffer(int((fb * fb).sum()))
            try:
                (_, _, res), (out, rcol) = _overlapped(
                    lambda: self._ctx.density2d_batch(specs, out=buf),
                    lambda: self._wrap_2d(pairs, specs, buf, offsets, conts, cache=not kwargs))
            except Exception:
                for pr in pairs:  # nothing of a failed batch stays in the cache
                    self._density2D.pop(tuple(pr), None)
                raise
            self._records_2d(pairs, rcol, res, cache=not kwargs)
            return out
        else:
            buf, offsets, res = self._ctx.density2d_batch(specs, out=_out)
        return self._finish_2d(pairs, specs, buf, offsets, res, conts, lbuf=lbuf, masks=masks,
                               cache=not kwargs and masks is None and lbuf is None)

    def _finish_2d(self, pairs, specs, buf, offsets, res, conts, lbuf=None, masks=None, cache=True):
        """Density2D objects (+ the reference's warnings / errors) from the grids of a 2D batch; `buf` holds density i at
        buf[offsets[i]:][:G*G]"""
        out, rcol = self._wrap_2d(pairs, specs, buf, offsets, conts, lbuf=lbuf, masks=masks, cache=cache)
        self._records_2d(pairs, rcol, res, cache=cache)
        return out

    def _wrap_2d(self, pairs, specs, buf, offsets, conts, lbuf=None, masks=None, cache=True):
        """The Density2D objects of a batch over (views of) its result buffer.  Needs nothing the device computes, so it
        can run while the library call is still in flight; the result records are attached by _records_2d afterwards.
        Returns (objects, the shared column dict their `_gdk` records read from)."""
        out = []
        names = self.paramNames.names
        # plain Python scalars of the spec columns once per batch (a field access per pair costs microseconds)
        col = {k: specs[k].tolist() for k in ("fine_bins", "xbinmin", "xbinmax", "ybinmin", "ybinmax", "bw_mode")}
        rcol = {"bw_mode": col["bw_mode"], "fine_bins": col["fine_bins"]}
        offsets = [int(o) for o in offsets]
        conts = list(conts)
        G0 = col["fine_bins"][0] if len(pairs) else 0
        grids = None
        if cache:
            buf = buf.view()
            buf.flags.writeable = False  # cache entries are read-only views into the batch buffer, see below
        if len(pairs) and all(g == G0 for g in col["fine_bins"]) and offsets == list(range(offsets[0], offsets[0] + len(pairs) * G0 * G0, G0 * G0)):
            grids = buf[offsets[0]: offsets[0] + len(pairs) * G0 * G0].reshape(len(pairs), G0, G0)  # one view, indexed per pair
        for n, (j, j2) in enumerate(pairs):
            parx, pary = names[j], names[j2]
            G = col["fine_bins"][n]
            off = offsets[n]
            d = Density2D.on_linspace((col["xbinmin"][n], col["xbinmax"][n], G), (col["ybinmin"][n], col["ybinmax"][n], G),
                                      grids[n] if grids is not None else buf[off: off + G * G].reshape(G, G),
                                      [(parx.range_min, parx.range_max), (pary.range_min, pary.range_max)])
            if masks is not None:  # bool_mask, mcsamples.py:1917, 1986
                w = len(masks[n]) - G
                w //= 2
                d.mask = np.asarray(masks[n][w: w + G, w: w + G] < 1e-8)
            d._gdk = _BatchRecord(rcol, n, conts)
            if lbuf is not None:
                d._likes2d = lbuf[off: off + G * G].reshape(G, G)
            if cache:
                # the cache entry is a READ-ONLY view into the batch buffer (no second copy of a gigabyte of grids);
                # get2DDensity / get2DDensityGridData hand out private copies of it (_cached_2d), as the reference
                # returns a fresh grid per call and callers normalise in place
                self._density2D[(j, j2)] = d
            out.append(d)
        return out, rcol

    def _records_2d(self, pairs, rcol, res, cache=True):
        """attach the library's result records to the objects of _wrap_2d and raise / warn as the reference does"""
        rcol.update(_abi.results2d_columns(res))
        bad = _abi.ST_BIAS_NEG | _abi.ST_BW_FALLBACK | _abi.ST_SMALL_SMOOTH | _abi.ST_ZERO_MAX
        names = self.paramNames.names
        try:
            for n, status in enumerate(rcol["status"]):
                if not status & bad:
                    continue
                parx, pary = names[pairs[n][0]], names[pairs[n][1]]
                if status & _abi.ST_BIAS_NEG:
                    raise Exception("bias not positive definite")  # kde_bandwidth.py:230-231 (propagates in the reference)
                if status & _abi.ST_BW_FALLBACK:
                    msg = "2D kernel density bandwidth optimizer failed for %s, %s. Using fallback width" % (parx.name, pary.name)
                    if self.raise_on_bandwidth_errors:
                        raise BandwidthError(msg)
                    log.warning(msg)
                if status & _abi.ST_SMALL_SMOOTH:
                    log.warning("fine_bins_2D not large enough for optimal density: %s, %s", parx.name, pary.name)
                if status & _abi.ST_ZERO_MAX:
                    raise DensitiesError("no samples in bin")
        except Exception:
            if cache:  # nothing of a failed batch stays behind
                for pr in pairs:
                    self._density2D.pop(tuple(pr), None)
            raise

    # ------------------------------------------------------------------ marginalised limits
    def setMargeLimits(self, params=None, max_frac_twotail=None):
        """Batched form of _setDensitiesandMarge1D (mcsamples.py:2442-2458): the 1D densities that are not cached yet
        in one device batch, every order statistic the limit logic can ask for (4 per contour) in ONE device quantile
        call, then the scalar logic of _setMargeLimits per parameter (getdist_b200/limits.py).  Sets `par.limits`
        (list of ParamLimit, one per contour) and returns {name: limits}."""
        from .limits import limit_fractions, marge_limits

        if self.needs_update:
            self.updateBaseStatistics()
        idx = list(range(self.n)) if params is None else [self._parAndNumber(p)[0] for p in params]
        if any(j is None for j in idx):
            raise ParamError("unknown parameter")
        missing = [j for j in idx if self.paramNames.names[j].name not in self.density1D]
        if missing:
            self._densities_1d(missing)
        mft = self.max_frac_twotail if max_frac_twotail is None else max_frac_twotail
        keys = limit_fractions(self.contours)
        out = {}
        for k0 in range(0, len(keys), 16):  # at most 16 target fractions per device call
            part = keys[k0:k0 + 16]
            fr = np.array([(1 - lf) if up else lf for lf, up in part])
            q = self._ctx.weighted_quantiles(idx, fr)
            for row, j in zip(q, idx):
                out.setdefault(j, {}).update(dict(zip(part, row)))
        res = {}
        for j in idx:
            par = self.paramNames.names[j]
            table = out[j]
            par.limits = marge_limits(self.density1D[par.name], par, self.contours, mft, lambda lf, up: table[(lf, up)],
                                      force_twotail=self.force_twotail,
                                      credible_interval_threshold=self.credible_interval_threshold)
            res[par.name] = par.limits
        return res

    def _setMargeLimits(self, par, paramConfid=None, max_frac_twotail=None, density1D=None):
        """mcsamples.py:2460-2531 for one parameter (paramConfid is not needed: the order statistics come from the
        device).  A density passed in replaces the cached one for this call only."""
        from .limits import limit_fractions, marge_limits

        j, par = self._parAndNumber(par if not isinstance(par, ParamInfo) else par.name)
        if density1D is None:
            density1D = self.get1DDensity(par.name)
        mft = self.max_frac_twotail if max_frac_twotail is None else max_frac_twotail
        keys = limit_fractions(self.contours)
        table = {}
        for k0 in range(0, len(keys), 16):  # at most 16 target fractions per device call
            part = keys[k0:k0 + 16]
            fr = np.array([(1 - lf) if up else lf for lf, up in part])
            table.update(zip(part, self._ctx.weighted_quantiles([j], fr)[0]))
        par.limits = marge_limits(density1D, par, self.contours, mft, lambda lf, up: table[(lf, up)],
                                  force_twotail=self.force_twotail, credible_interval_threshold=self.credible_interval_threshold)
        return par.limits

    # ------------------------------------------------------------------ raw ND densities
    def getRawNDDensity(self, xs, normalized=False, **kwargs):
        """mcsamples.py:2081-2096."""
        if self.needs_update:
            self.updateBaseStatistics()
        density = self.getRawNDDensityGridData(xs, get_density=True, **kwargs)
        if normalized:
            density.normalize(in_place=True)
        return density

    def getRawNDDensityGridData(self, js, writeDataToFile=False, num_plot_contours=None, get_density=False, meanlikes=False,
                                maxlikes=False, **kwargs):
        """mcsamples.py:2098-2235: unsmoothed ND marginalised density.  The N-sized part (_binSamples per axis +
        _makeNDhist, and the mean / profile likelihood histograms) is one device sweep per histogram (gdk_histnd);
        the raw edge mask, the normalisation and the contour levels are grid-sized host arithmetic."""
        if writeDataToFile:
            raise NotImplementedError("plot-data files are written by the reference (file IO is out of scope)")
        if self.needs_update:
            self.updateBaseStatistics()
        jv, parv = zip(*[self._parAndNumber(j) for j in js])
        if None in jv:
            return None
        ndim = len(jv)
        self._ensure_param_ranges(list(jv))
        bco = kwargs.get("boundary_correction_order", self.boundary_correction_order)
        has_prior = any(par.has_limits for par in parv)
        nbinsND = int(kwargs.get("num_bins_ND", self.num_bins_ND))
        geo = [self._bin_geometry(par, nbinsND) for par in parv]
        lo, hi = [g[0] for g in geo], [g[1] for g in geo]
        nb = [nbinsND] * ndim
        if meanlikes or maxlikes:
            self._ensure_loglikes()
        binsND = self._ctx.histnd(jv, nb, lo, hi, 0)
        if has_prior and bco >= 0:
            # _setRawEdgeMaskND (mcsamples.py:2012-2032): half weight on the boundary bins of every hard prior
            prior_mask = np.ones(binsND.shape)
            for ax, par in enumerate(parv[::-1]):
                sl = [slice(None)] * ndim
                if par.has_limits_bot:
                    sl[ax] = 0
                    prior_mask[tuple(sl)] /= 2
                if par.has_limits_top:
                    sl[ax] = binsND.shape[ax] - 1
                    prior_mask[tuple(sl)] /= 2
            binsND = binsND / prior_mask
        xv = [np.linspace(lo[i], hi[i], nb[i]) for i in range(ndim)]
        views = [(par.range_min, par.range_max) for par in parv]
        density = DensityND(xv, binsND, view_ranges=views)
        density.normalize("max", in_place=True)
        if get_density:
            return density
        ncontours = len(self.contours)
        if num_plot_contours:
            ncontours = min(num_plot_contours, ncontours)
        contours = self.contours[:ncontours]
        density.contours = density.getContourLevels(contours)
        if meanlikes:
            likes = self._ctx.histnd(jv, nb, lo, hi, 1)
            density.likes = likes / np.max(likes)
        else:
            density.likes = None
        if maxlikes:
            from .densities import getContourLevels

            density.maxlikes = self._ctx.histnd(jv, nb, lo, hi, 2)
            # getImportContourLevels(binNDmaxlikes, contours, half_edge=False) == getContourLevels(..., half_edge=False)
            density.maxcontours = getContourLevels(density.maxlikes, contours, half_edge=False)
        else:
            density.maxlikes = None
        return density

    # ------------------------------------------------------------------ convergence tests
    def getFractionIndices(self, weights=None, n=2):
        """mcsamples.py:668-680 on the stored weights: rows that split the total weight into n equal parts."""
        if weights is not None and weights is not self.weights:
            raise NotImplementedError("getFractionIndices works on the stored weights")
        if self.needs_update:
            self.updateBaseStatistics()
        rows = self._ctx.weight_fraction_rows(np.linspace(0, 1, n, endpoint=False))
        return np.append(rows, self.numrows)

    def getMeanVarTest(self):
        """'MeanVar' of getConvergeTests (mcsamples.py:964-989): per parameter sqrt(var(chain mean) / mean(chain var)),
        from the per-chain moments of the fused device reduction."""
        if self.chain_offsets is None:
            raise WeightedSampleError("Samples were not combined from separate chains")
        if self.needs_update:
            self.updateBaseStatistics()
        m = self._mom
        nch = m["chain_means"].shape[0]
        between = np.sum((m["chain_means"] - self.means) ** 2, axis=0) / (nch - 1)
        # in-chain variance: sum over chains of sum w (x - chain mean)^2, over the total weight
        within = np.zeros(self.n)
        for c in range(nch):
            within += np.diagonal(m["chain_covs"][c]) * m["chain_norms"][c]
        within /= self.norm
        return np.sqrt(between / within)

    def getSplitTests(self, test_confidence=0.95):
        """'SplitTest' of getConvergeTests (mcsamples.py:1003-1034): rms change of the upper / lower quantile, in units
        of the standard deviation, when the samples are split into 2 .. max_split_tests sets of equal weight.  Returns
        (nparam, max_split_tests - 1, 2).  Every order statistic is a device call (gdk_weighted_quantiles[_range]),
        batched over all parameters."""
        if self.needs_update:
            self.updateBaseStatistics()
        limits = np.array([1 - (1 - test_confidence) / 2, (1 - test_confidence) / 2])
        idx = list(range(self.n))
        confids = self._ctx.weighted_quantiles(idx, limits)
        out = np.zeros((self.n, self.max_split_tests - 1, 2))
        for ix in range(self.max_split_tests - 1):
            split_n = 2 + ix
            frac = self.getFractionIndices(None, split_n)
            for f1, f2 in zip(frac[:-1], frac[1:]):
                out[:, ix, :] += (self._ctx.weighted_quantiles_range(idx, limits, f1, f2) - confids) ** 2
            out[:, ix, :] = np.sqrt(out[:, ix, :] / split_n) / self.sddev[:, None]
        return out

    def getConvergeTests(self, test_confidence=0.95, writeDataToFile=False, what=("MeanVar", "GelmanRubin", "SplitTest"),
                         filename=None, feedback=False):
        """mcsamples.py:905-1034: the MeanVar, GelmanRubin and SplitTest sections, same text as the reference.
        'RafteryLewis' and 'CorrLengths' (thinned binary chains, per-chain full-length autocorrelations) are not on
        the device path."""
        other = [w for w in what if w not in ("MeanVar", "GelmanRubin", "SplitTest")]
        if other:
            raise NotImplementedError("convergence tests %s are not on the device path" % other)
        if writeDataToFile:
            raise NotImplementedError("file output stays with the reference")
        if self.needs_update:
            self.updateBaseStatistics()
        lines = ""
        nparam = self.n
        nch = 1 if self.chain_offsets is None else len(self.chain_offsets) - 1
        parForm = self.paramNames.parFormat()
        parNames = [parForm % self.paramNames.names[j].name for j in range(nparam)]
        if nch > 1 and "MeanVar" in what:
            lines += "\n"
            lines += "mean convergence stats using remaining chains\n"
            lines += "param sqrt(var(chain mean)/mean(chain var))\n"
            lines += "\n"
            mv = self.getMeanVarTest()
            for j in range(nparam):
                lines += parNames[j] + f"{mv[j]:10.4f}  {self.paramNames.names[j].label}\n"
            lines += "\n"
        nparamMC = self.paramNames.numNonDerived()
        if nch > 1 and nparamMC > 0 and "GelmanRubin" in what:
            D = self.getGelmanRubinEigenvalues()
            if D is not None:
                self.GelmanRubin = np.max(D)
                lines += "var(mean)/mean(var) for eigenvalues of covariance of y of orthonormalized parameters\n"
                for jj, Di in enumerate(D):
                    lines += "%3i%13.5f\n" % (jj + 1, Di)
                GRSummary = " var(mean)/mean(var), remaining chains, worst e-value: R-1 = %13.5F" % self.GelmanRubin
            else:
                self.GelmanRubin = None
                GRSummary = "Gelman-Rubin covariance not invertible (parameter not moved?)"
                log.warning(GRSummary)
            if feedback:
                print(GRSummary)
            lines += "\n"
        if "SplitTest" in what:
            lines += "Split tests: rms_n([delta(upper/lower quantile)]/sd) n={2,3,4}, limit=%.0f%%:\n" % (100 * self.converge_test_limit)
            lines += "i.e. mean sample splitting change in the quantiles in units of the st. dev.\n"
            lines += "\n"
            st = self.getSplitTests(test_confidence)
            for j in range(nparam):
                for endb, typestr in enumerate(["upper", "lower"]):
                    lines += parNames[j]
                    for ix in range(self.max_split_tests - 1):
                        lines += "%9.4f" % (st[j, ix, endb])
                    lines += " %s\n" % typestr
            lines += "\n"
        if feedback:
            print(lines)
        return lines

    # ------------------------------------------------------------------ batched driver
    def triangle_pairs(self, params=None):
        idx = list(range(self.n)) if params is None else [self._parAndNumber(p)[0] for p in params]
        if any(i is None for i in idx):
            raise ParamError("unknown parameter")
        return idx, [(idx[i], idx[k]) for i in range(len(idx)) for k in range(i + 1, len(idx))]

    def invalidate_density_caches(self):
        """Forget per-parameter ranges and cached densities (as updateBaseStatistics does) without touching the
        resident samples or the moments: the next density call recomputes quantiles, ranges and grids."""
        self.density1D = {}
        self._density2D = {}
        for par in self.paramNames.names:
            par._ranges_ready = False
            par.N_eff_kde = None
        self._initLimits()

    def prefetch_triangle(self, params=None, do_1d=True, do_2d=True, root=None):
        """Compute every 1D and (lower-triangle) 2D density of a triangle plot in batched launches and seed
        the caches that get1DDensity / get2DDensity consult.  Pair (x, y) = (params[i], params[k]) for i < k,
        as getdist.plots.triangle_plot requests them (x = column parameter, y = row parameter).
        Returns the cache entries themselves: the 2D grids are read-only views into one (pinned) result buffer.
        With a process group (MCSamples(process_group=...)): every rank computes its share; root=None leaves every rank
        with every density, root=r only rank r (the others return empty lists)."""
        if self.needs_update:
            self.updateBaseStatistics()
        idx = list(range(self.n)) if params is None else [self._parAndNumber(p)[0] for p in params]
        if any(i is None for i in idx):
            raise ParamError("unknown parameter in prefetch_triangle")
        pg = getattr(self, "process_group", None)
        if pg is not None and pg.world > 1:
            from .parallel import prefetch_triangle_group

            return prefetch_triangle_group(self, pg, idx, do_1d, do_2d, root=root)
        import time as _time

        t0 = _time.perf_counter()
        d1 = self._densities_1d(idx) if do_1d else []
        t1 = _time.perf_counter()
        pairs = [(idx[i], idx[k]) for i in range(len(idx)) for k in range(i + 1, len(idx))]
        d2 = self._densities_2d(pairs) if (do_2d and pairs) else []
        t2 = _time.perf_counter()
        wl = self._ctx.wall_ms() if hasattr(self._ctx, "wall_ms") else {}
        # host wall clock of the call: the 1D and 2D batches, of which the time inside the library's 2D call
        self.last_prefetch_ms = dict(d1=(t1 - t0) * 1e3, d2=(t2 - t1) * 1e3, lib_2d=wl.get("call_2d"), lib_1d=wl.get("call_1d"),
                                     lib_quantiles=wl.get("call_quantiles"))
        return d1, d2
