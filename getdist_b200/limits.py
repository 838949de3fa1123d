"""Marginalised 1D parameter limits: the scalar logic of MCSamples._setMargeLimits
(mcsamples.py:2460-2531 of the reference) on top of a 1D density grid, its spline-refined equal-density limits
(Density1D.getLimits) and exact weighted order statistics (the device quantile kernels).  Host-side and grid-sized;
kept free of device calls so that it is tested on the CPU against limits produced by the unmodified reference."""
import math


class ParamLimit:
    """One marginalised limit (types.py:652-691 of the reference): lower / upper and the kind of limit."""

    def __init__(self, minmax, tag="two"):
        self.lower, self.upper = minmax[0], minmax[1]
        self.twotail = tag == "two"
        self.onetail_upper = tag == ">"
        self.onetail_lower = tag == "<"

    def limitTag(self):
        return "two" if self.twotail else (">" if self.onetail_upper else ("<" if self.onetail_lower else "none"))

    def __repr__(self):
        return "ParamLimit(%r, %r, %r)" % (self.lower, self.upper, self.limitTag())


def limit_fractions(contours):
    """the probability fractions whose order statistics marge_limits may ask for, as (limfrac, upper) keys"""
    keys = []
    for c in contours:
        lf = 1 - c
        keys += [(lf, False), (lf, True), (lf / 2, False), (lf / 2, True)]
    return keys


def marge_limits(density, par, contours, max_frac_twotail, confidence, force_twotail=False,
                 credible_interval_threshold=0.05):
    """List of ParamLimit, one per contour.  `confidence(limfrac, upper)` returns the exact weighted order statistic
    (chains.py:814-838); `par` carries has_limits_bot/top and range_min/max as left by _initParam."""
    limits = []
    grids = None
    for k, contour in enumerate(contours):
        # a bounded side whose end bin is still high relative to the peak counts as "no limit on that side"
        open_bot = par.has_limits_bot and not force_twotail and density.P[0] > max_frac_twotail[k]
        open_top = par.has_limits_top and not force_twotail and density.P[-1] > max_frac_twotail[k]
        if open_bot and open_top:
            limits.append(ParamLimit([par.range_min, par.range_max], "none"))
            continue
        if grids is None:
            grids = density.initLimitGrids()
        lo, hi, open_bot, open_top = density.getLimits(contour, grids)
        limfrac = 1 - contour
        q_lo = q_hi = None
        if open_bot:
            lo = par.range_min  # pinned to the end of the prior range
        elif open_top:
            lo = confidence(limfrac, False)  # one-tail limit
        else:
            q_lo = confidence(limfrac / 2, False)
        if open_top:
            hi = par.range_max
        elif open_bot:
            hi = confidence(limfrac, True)
        else:
            q_hi = confidence(limfrac / 2, True)
        if not open_bot and not open_top:
            # two tails: equal-tail quantiles unless the density differs a lot between them (then the credible interval)
            if math.fabs(density.Prob(q_hi) - density.Prob(q_lo)) < credible_interval_threshold:
                lo, hi = q_lo, q_hi
        tag = "none" if (open_bot and open_top) else (">" if open_bot else ("<" if open_top else "two"))
        limits.append(ParamLimit([lo, hi], tag))
    return limits
