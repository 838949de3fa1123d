"""Drop-in adapter for the reference: a subclass of ``getdist.mcsamples.MCSamples`` whose hot-path methods run on the
B200 through libgdk.so, so that ``getdist.plots`` / the GUIs / the CLI are unchanged callers (SURVEY.md s8b;
``plots.py:594-645, 1116, 2613``).

    from getdist_b200.reference_backend import MCSamples      # instead of: from getdist import MCSamples
    s = MCSamples(samples=X, weights=w, names=[...], ranges={...})
    g = getdist.plots.get_subplot_plotter(); g.triangle_plot(s, params)       # the plot loop is batched transparently

Overridden (everything else -- file IO, settings/.ini handling, PCA, tables, likeStats -- is inherited):

    setMeans / getVars / cov            chains.py:373-412, 709-733     (one fused device reduction)
    getGelmanRubinEigenvalues           chains.py:1446-1474            (per-chain moments from the same reduction)
    confidence                          chains.py:814-838              (exact device order statistics)
    get1DDensityGridData                mcsamples.py:1517-1686
    get2DDensityGridData                mcsamples.py:1748-2010
    getRawNDDensityGridData             mcsamples.py:2098-2235
    _setDensitiesandMarge1D             mcsamples.py:2442-2458         (batched 1D densities + limits)

The device object (``getdist_b200.MCSamples``, the host-side planner over the C-ABI) is created lazily, is rebuilt when
the samples / weights / chain boundaries change, and is re-synchronised with the analysis settings and the hard ranges
of this object whenever they change (``updateSettings``, ``setRanges``, direct attribute assignment).  Requests that
are not on the device path (``chainlist`` of foreign chain objects, ``writeDataToFile``, signed mean-likelihood
weights, ``range_ND_contour`` with likeStats) are routed to the inherited reference method -- explicitly, per call.

This module needs the reference importable (``import getdist``); the package itself does not.
"""
import copy as _copy

import numpy as np

try:
    from getdist import mcsamples as _ref
    from getdist import types as _types
    from getdist.densities import Density1D as _RefDensity1D
    from getdist.densities import Density2D as _RefDensity2D
    from getdist.densities import DensityND as _RefDensityND
except ImportError as e:  # pragma: no cover
    raise ImportError("getdist_b200.reference_backend adapts the reference's MCSamples class and needs `getdist` "
                      "importable; use getdist_b200.MCSamples for the standalone mirror") from e

from . import _abi
from .mcsamples import MCSamples as _GPU

# analysis settings the device planner reads (mcsamples.py:380-440 of the reference)
_SETTINGS = ("fine_bins", "fine_bins_2D", "smooth_scale_1D", "smooth_scale_2D", "boundary_correction_order",
             "mult_bias_correction_order", "max_corr_2D", "range_confidence", "num_bins", "num_bins_2D", "num_bins_ND",
             "range_ND_contour", "credible_interval_threshold", "use_effective_samples_2D", "converge_test_limit")
_FLAGS = ("raise_on_bandwidth_errors", "force_twotail", "no_warning_params", "no_warning_chi2_params",
          "shade_likes_is_mean_loglikes", "max_split_tests")
_PAR_ATTRS = ("err", "mean", "param_min", "param_max", "range_min", "range_max", "sigma_range", "has_limits_bot",
              "has_limits_top", "has_limits", "kde_h", "N_eff_kde")


class MCSamples(_ref.MCSamples):
    """Drop-in replacement of getdist.MCSamples with the density / statistics hot path on the GPU."""

    gpu_device = 0

    # ------------------------------------------------------------------ device object
    def _gpu_state_key(self):
        rg = self.ranges
        names = self.paramNames.list()
        return (tuple((k, repr(getattr(self, k, None))) for k in _SETTINGS + _FLAGS), tuple(np.asarray(self.contours).tolist()),
                tuple((n, rg.getLower(n), rg.getUpper(n), n in getattr(rg, "periodic", ())) for n in names), tuple(names),
                self.sampler)

    def _gpu_ranges(self):
        rg = self.ranges
        per = getattr(rg, "periodic", ())
        out = {}
        for n in self.paramNames.list():
            lo, hi = rg.getLower(n), rg.getUpper(n)
            if lo is not None or hi is not None:
                out[n] = (lo, hi, True) if n in per else (lo, hi)
        return out

    def _gpu(self):
        """the device-backed planner for the current samples, settings and ranges"""
        g = self.__dict__.get("_gpu_obj")
        offs = None if self.chain_offsets is None else np.asarray(self.chain_offsets)
        if g is not None and (g.samples is not self.samples or g.weights is not self.weights or g.loglikes is not self.loglikes
                              or not np.array_equal(g.chain_offsets if g.chain_offsets is not None else (), offs if offs is not None else ())):
            g = None
        key = self._gpu_state_key()
        if g is None:
            settings = {k: getattr(self, k) for k in _SETTINGS if hasattr(self, k)}
            settings["contours"] = np.asarray(self.contours)
            g = _GPU(samples=self.samples, weights=self.weights, loglikes=self.loglikes, names=self.paramNames.list(),
                     labels=[p.label for p in self.paramNames.names], ranges=self._gpu_ranges(), sampler=self.sampler,
                     settings=settings, chain_offsets=offs, device=self.gpu_device)
            for par, gp in zip(self.paramNames.names, g.paramNames.names):
                gp.isDerived = par.isDerived
            for k in _FLAGS:
                if hasattr(self, k):
                    setattr(g, k, getattr(self, k))
            self.__dict__["_gpu_obj"] = g
            self.__dict__["_gpu_key"] = key
        elif self.__dict__.get("_gpu_key") != key:
            # settings / ranges changed on the reference object since the last call: re-sync, forget the derived caches
            from .mcsamples import ParamBounds

            settings = {k: getattr(self, k) for k in _SETTINGS if hasattr(self, k)}
            settings["contours"] = np.asarray(self.contours)
            g.updateSettings(settings, doUpdate=False)
            for k in _FLAGS:
                if hasattr(self, k):
                    setattr(g, k, getattr(self, k))
            g.sampler = self.sampler
            g.ranges = ParamBounds(self._gpu_ranges())
            g.invalidate_density_caches()
            self.__dict__["_gpu_key"] = key
        return g

    def _weightsChanged(self):  # chains.py:310-323: the device copy goes with the host statistics
        super()._weightsChanged()
        self.__dict__.pop("_gpu_obj", None)

    def __getstate__(self):  # stay picklable / deep-copyable (mcsamples.py:125, 316): the device object is rebuilt lazily
        d = self.__dict__.copy()
        d.pop("_gpu_obj", None)
        d.pop("_gpu_key", None)
        return d

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__dict__.update({k: _copy.deepcopy(v, memo) for k, v in self.__getstate__().items()})
        return new

    def _sync_param(self, par, gp):
        for a in _PAR_ATTRS:
            if hasattr(gp, a):
                setattr(par, a, getattr(gp, a))

    # ------------------------------------------------------------------ statistics
    def setMeans(self):
        g = self._gpu()
        self.means = g.getMeans()
        if self.loglikes is not None:
            g._ensure_loglikes()
            self.mean_loglike = g.mean_loglike
        else:
            self.mean_loglike = None
        return self.means

    def getVars(self):
        g = self._gpu()
        if self.means is None:
            self.setMeans()
        self.vars = g.getVars()
        self.sddev = np.sqrt(self.vars)
        return self.vars

    def cov(self, pars=None, where=None):
        if pars is None and where is None:
            return self._gpu().getCov().copy()
        return super().cov(pars, where)

    def getGelmanRubinEigenvalues(self, nparam=None, chainlist=None):
        if chainlist is not None or self.chain_offsets is None:
            return super().getGelmanRubinEigenvalues(nparam, chainlist)  # foreign chain objects / separate chain files
        return self._gpu().getGelmanRubinEigenvalues(nparam or self.paramNames.numNonDerived())

    def confidence(self, paramVec, limfrac, upper=False, start=0, end=None, weights=None):
        if isinstance(paramVec, (int, np.integer)) and 0 <= paramVec < self.n and weights is None:
            return self._gpu().confidence(int(paramVec), limfrac, upper, start, end)
        return super().confidence(paramVec, limfrac, upper, start, end, weights)

    # ------------------------------------------------------------------ densities
    def get1DDensityGridData(self, j, paramConfid=None, meanlikes=False, **kwargs):
        if meanlikes and self.shade_likes_is_mean_loglikes:
            return super().get1DDensityGridData(j, paramConfid, meanlikes, **kwargs)  # signed weights: not accelerated
        if self.needs_update:
            self.updateBaseStatistics()
        jj, par = self._parAndNumber(j)
        if jj is None:
            return None
        g = self._gpu()
        d = g.get1DDensity(jj) if not (meanlikes or kwargs) else g.get1DDensityGridData(jj, meanlikes=meanlikes, **kwargs)
        self._sync_param(par, g.paramNames.names[jj])
        out = _RefDensity1D(d.x, d.P.copy(), view_ranges=list(d.view_ranges))
        out.likes = None if d.likes is None else d.likes.copy()
        if not kwargs:
            self.density1D[par.name] = out  # mcsamples.py:1669-1670
        return out

    def get2DDensityGridData(self, j, j2, num_plot_contours=None, get_density=False, meanlikes=False, mask_function=None,
                             **kwargs):
        if self.needs_update:
            self.updateBaseStatistics()
        jx, parx = self._parAndNumber(j)
        jy, pary = self._parAndNumber(j2)
        if jx is None or jy is None:
            return None
        if mask_function is not None and (getattr(parx, "periodic", False) or getattr(pary, "periodic", False)):
            return super().get2DDensityGridData(j, j2, num_plot_contours, get_density, meanlikes, mask_function, **kwargs)
        g = self._gpu()
        d = g.get2DDensityGridData(jx, jy, num_plot_contours, get_density, meanlikes, mask_function, **kwargs)
        self._sync_param(parx, g.paramNames.names[jx])
        self._sync_param(pary, g.paramNames.names[jy])
        out = _RefDensity2D(d.x, d.y, d.P, view_ranges=[tuple(v) for v in d.view_ranges], mask=getattr(d, "mask", None))
        if not get_density:
            out.contours = d.contours
            out.likes = d.likes
        return out

    def getRawNDDensityGridData(self, js, writeDataToFile=False, num_plot_contours=None, get_density=False, meanlikes=False,
                                maxlikes=False, **kwargs):
        if writeDataToFile:
            return super().getRawNDDensityGridData(js, writeDataToFile, num_plot_contours, get_density, meanlikes, maxlikes, **kwargs)
        if self.needs_update:
            self.updateBaseStatistics()
        jv = [self._parAndNumber(j)[0] for j in js]
        if None in jv:
            return None
        d = self._gpu().getRawNDDensityGridData(jv, False, num_plot_contours, get_density, meanlikes, maxlikes, **kwargs)
        out = _RefDensityND(d.xs, d.P, view_ranges=d.view_ranges)
        if not get_density:
            out.contours, out.likes, out.maxlikes = d.contours, d.likes, d.maxlikes
            if maxlikes:
                out.maxcontours = d.maxcontours
        return out

    def _setDensitiesandMarge1D(self, max_frac_twotail=None, meanlikes=False):
        """getMargeStats' 1D densities + limits (mcsamples.py:2442-2458), batched on the device"""
        if self.done_1Dbins:
            return
        if meanlikes or (self.range_ND_contour >= 0 and self.likeStats):
            return super()._setDensitiesandMarge1D(max_frac_twotail, meanlikes)
        g = self._gpu()
        for name, lims in g.setMargeLimits(None, max_frac_twotail).items():
            par = self.paramNames.parWithName(name)
            self._sync_param(par, g.paramNames.names[g.index[name]])
            par.limits = [_types.ParamLimit([lim.lower, lim.upper], lim.limitTag()) for lim in lims]
            d = g.density1D[name]
            self.density1D[name] = _RefDensity1D(d.x, d.P.copy(), view_ranges=list(d.view_ranges))
        self.done_1Dbins = True

    # ------------------------------------------------------------------ batched driver
    def prefetch_triangle(self, params=None, **kw):
        """all 1D and 2D densities of a triangle plot in batched launches (optional: the plot loop batches by itself)"""
        if self.needs_update:
            self.updateBaseStatistics()
        idx = None if params is None else [self._parAndNumber(p)[0] for p in params]
        return self._gpu().prefetch_triangle(idx, **kw)


def loadMCSamples(*args, **kwargs):
    """getdist.loadMCSamples returning the adapter class: the reference reads the chain files, the result is re-typed"""
    s = _ref.loadMCSamples(*args, **kwargs)
    s.__class__ = MCSamples
    return s
