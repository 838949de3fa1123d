"""Return types of the density calls -- host float64 grids with the attributes getdist.plots reads
(`.x/.y/.P/.contours/.likes/.mask/.view_ranges/.bounds()`), mirroring getdist.densities
(densities.py:59-280 of the reference) so that plotting code receives the same kind of object."""
import numpy as np


class DensitiesError(Exception):
    pass


def getContourLevels(inbins, contours=(0.68, 0.95), missing_norm=0, half_edge=True):
    """Density levels enclosing the given probability fractions (reference densities.py:19-56): grid cells are
    visited in order of increasing density, their (edge-halved) masses accumulated, and the level is interpolated
    between the two neighbouring sorted densities where the running mass passes (1 - contour) of the total.
    Host-side numpy on a returned grid; the batched path computes the same levels on the device (k_contours2d)."""
    fracs = np.atleast_1d(contours)
    mass = np.array(inbins, dtype=np.float64, copy=True)
    if half_edge:  # trapezoid weights: boundary cells count half along every axis
        for ax in range(mass.ndim):
            edge = [slice(None)] * mass.ndim
            for end in (0, -1):
                edge[ax] = end
                mass[tuple(edge)] *= 0.5
    rank = np.argsort(inbins, axis=None)
    dens = mass.reshape(-1)[rank]
    run = np.cumsum(dens)
    want = (1 - np.asarray(fracs)) * mass.sum() - missing_norm
    levels = np.zeros(len(fracs))
    for k, (i, t) in enumerate(zip(np.searchsorted(run, want), want)):
        if i == 0:
            raise DensitiesError("Contour level outside plotted ranges")
        back = (run[i] - t) / (run[i] - run[i - 1])  # how far back towards the previous sorted cell
        levels[k] = dens[i] * (1 - back) + back * dens[i - 1]
    return levels


class GridDensity:
    def normalize(self, by="integral", in_place=False):
        if by == "integral":
            norm = self.norm_integral()
        elif by == "max":
            norm = np.max(self.P)
            if norm == 0:
                raise DensitiesError("no samples in bin")
        else:
            raise DensitiesError("Density: unknown normalization")
        if in_place:
            self.P /= norm
        else:
            self.setP(self.P / norm)
        self.spl = None
        return self

    def setP(self, P=None):
        if P is not None:
            for size, ax in zip(P.shape, self.axes):
                if size != ax.size:
                    raise DensitiesError("Array size mismatch in Density arrays: P %s, axis %s" % (size, ax.size))
            self.P = P
        else:
            self.P = np.zeros([ax.size for ax in self.axes])
        self.spl = None

    def bounds(self):
        """bounds in the order x, y, z ... (the axes are stored slowest first, reference densities.py:111-121)"""
        if self.view_ranges is not None:
            return self.view_ranges
        return [(ax[0], ax[-1]) for ax in self.axes][::-1]

    def getContourLevels(self, contours=(0.68, 0.95)):
        return getContourLevels(self.P, contours)


class Density1D(GridDensity):
    def __init__(self, x, P=None, view_ranges=None):
        self.n = x.size
        self.axes = [x]
        self.x = x
        self.view_ranges = view_ranges
        self.spacing = x[1] - x[0]
        self.likes = None
        self.setP(P)

    def bounds(self):
        if self.view_ranges is not None:
            return self.view_ranges
        return self.x[0], self.x[-1]

    def integrate(self, P):
        return ((P[0] + P[-1]) / 2 + np.sum(P[1:-1])) * self.spacing

    def norm_integral(self):
        return self.integrate(self.P)

    def Prob(self, x, derivative=0):
        from scipy.interpolate import splev, splrep

        if self.spl is None:
            self.spl = splrep(self.x, self.P, s=0)
        if isinstance(x, (np.ndarray, list, tuple)):
            return splev(x, self.spl, derivative, ext=1)
        return splev([x], self.spl, derivative, ext=1)[0]

    __call__ = Prob

    # ---- equal-density credible limits (densities.py:186-248 of the reference): host-side, grid-sized ----------
    def initLimitGrids(self, factor=None):
        """Spline-refined copy of the density (factor x finer), its total and the ascending cumulative sum of its
        sorted values: the tables getLimits() searches (reference densities.py:186-205)."""
        from scipy.interpolate import splev, splrep

        if self.spl is None:
            self.spl = splrep(self.x, self.P, s=0)
        g = _LimitGrids()
        g.factor = max(2, 20000 // self.n) if factor is None else factor
        g.bign = (self.n - 1) * g.factor + 1
        g.grid = splev(self.x[0] + np.arange(g.bign) * self.spacing / g.factor, self.spl)
        g.norm = np.sum(g.grid) - 0.5 * self.P[-1] - 0.5 * self.P[0]
        g.sortgrid = np.sort(g.grid)
        g.cumsum = np.cumsum(g.sortgrid)
        return g

    def getLimits(self, p, interpGrid=None, accuracy_factor=None):
        """(min, max, has_min_boundary, has_max_boundary) of the equal-density interval holding probability p
        (tuple for a scalar p, list of tuples otherwise); a side whose end bin lies above the density level has no
        limit there (reference densities.py:207-248)."""
        g = interpGrid or self.initLimitGrids(accuracy_factor)
        fracs = np.atleast_1d(p)
        targets = (1 - fracs) * g.norm
        fine = self.spacing / g.factor
        out = []
        for ix, target in zip(np.searchsorted(g.cumsum, targets), targets):
            level = g.sortgrid[ix]
            if ix > 0:
                step = g.cumsum[ix] - g.cumsum[ix - 1]
                t = (g.cumsum[ix] - target) / step
                level = (1 - t) * level + t * g.sortgrid[ix + 1]
            at_bot = g.grid[0] >= level
            if at_bot:
                lo = self.x[0]
            else:
                i = int(np.argmax(g.grid > level))  # first fine bin above the level; interpolate the crossing
                lo = self.x[0] + (i - (g.grid[i] - level) / (g.grid[i] - g.grid[i - 1])) * fine
            at_top = g.grid[-1] >= level
            if at_top:
                hi = self.x[-1]
            else:
                i = g.bign - int(np.argmax(g.grid[::-1] > level)) - 1  # last fine bin above the level
                hi = self.x[0] + (i + (g.grid[i] - level) / (g.grid[i] - g.grid[i + 1])) * fine
            if fracs is not p:
                return lo, hi, at_bot, at_top
            out.append((lo, hi, at_bot, at_top))
        return out


class _LimitGrids:
    """tables built by Density1D.initLimitGrids"""

    factor = bign = grid = norm = sortgrid = cumsum = None


class Density2D(GridDensity):
    def __init__(self, x, y, P=None, view_ranges=None, mask=None):
        self._x = x
        self._y = y
        self.view_ranges = view_ranges
        self.mask = mask
        self.likes = None
        self.contours = None
        self.setP(P)

    @classmethod
    def on_linspace(cls, xr, yr, P, view_ranges=None):
        """Grid on x = linspace(*xr), y = linspace(*yr) (xr = (lo, hi, n)).  The axis vectors are only built when
        something reads them: the batched calls hand back thousands of these per launch."""
        d = cls.__new__(cls)
        d._x = d._y = None
        d._xr, d._yr = xr, yr
        d.view_ranges = view_ranges
        d.mask = d.likes = d.contours = d.spl = None
        if P.shape != (yr[2], xr[2]):
            raise DensitiesError("Array size mismatch in Density arrays: P %s, axes %s" % (P.shape, (yr[2], xr[2])))
        d.P = P
        return d

    @property
    def x(self):
        if self._x is None:
            self._x = np.linspace(*self._xr)
        return self._x

    @x.setter
    def x(self, v):
        self._x = v

    @property
    def y(self):
        if self._y is None:
            self._y = np.linspace(*self._yr)
        return self._y

    @y.setter
    def y(self, v):
        self._y = v

    @property
    def axes(self):
        return [self.y, self.x]

    @property
    def spacing(self):
        return (self.x[1] - self.x[0]) * (self.y[1] - self.y[0])

    def integrate(self, P):
        """trapezoid rule on the grid: corners weigh 1/4, edges 1/2 (reference densities.py:273-280)"""
        wy = np.ones(P.shape[0])
        wx = np.ones(P.shape[1])
        wy[[0, -1]] = 0.5
        wx[[0, -1]] = 0.5
        return float(wy @ P @ wx) * self.spacing

    def norm_integral(self):
        return self.integrate(self.P)

    def Prob(self, x, y, grid=False):
        from scipy.interpolate import RectBivariateSpline

        if self.spl is None:
            self.spl = RectBivariateSpline(self.x, self.y, self.P.T, s=0)
        return self.spl.ev(x, y) if not grid else self.spl(x, y)

    __call__ = Prob


class DensityND(GridDensity):
    """ND marginalised density on a regular grid (reference densities.py:304-381): `xs` lists the axis vectors in the
    order x, y, z ..., P is indexed slowest axis first (P[..., iy, ix])."""

    def __init__(self, xs, P=None, view_ranges=None):
        self.dim = len(xs)
        self.x = xs[0]
        if self.dim >= 2:
            self.y = xs[1]
        if self.dim >= 3:
            self.z = xs[2]
        self.xs = xs
        self.axes = xs[::-1]
        self.view_ranges = view_ranges
        self.spacing = 1.0
        for x in xs:
            self.spacing = self.spacing * (x[1] - x[0])
        self.likes = None
        self.maxlikes = None
        self.contours = None
        self.setP(P)

    def integrate(self, P):
        """every grid point weighs 2^-(number of axes on which it sits on the boundary); like the reference
        (densities.py:337-365) the sum is NOT multiplied by the cell volume"""
        P = np.asarray(P)
        w = np.ones(())
        for size in P.shape:
            e = np.ones(size)
            e[[0, -1]] = 0.5
            w = np.multiply.outer(w, e)
        return float(np.sum(P * w))

    def norm_integral(self):
        return self.integrate(self.P)

    def Prob(self, xs):
        from scipy.interpolate import LinearNDInterpolator

        if self.spl is None:
            self.spl = LinearNDInterpolator(self.xs, self.P.T, rescale=True)
        return self.spl(xs)

    __call__ = Prob
