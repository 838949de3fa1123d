// kde2d_core.cuh -- grid-sized bandwidth selection for one 2D density, executed by one cooperating group.
//
// Restates KernelOptimizer2D (kde_bandwidth.py:146-309): psi functionals as bilinear forms on the squared
// 2D-DCT (even orders) and on |FFT2|^2 (odd orders), the func2d / func2d_odd plug-in recursions, the
// Brent fixed point for t*, the closed-form h_x, h_y, and the AMISE minimisation with its accept rules;
// and the tail of getAutoBandwidth2D (mcsamples.py:1376-1419): rescaling to parameter units, de-rotation of
// the sheared kernel, bias-order rescale, fallback widths.
//
// Differences from the reference, both documented in DESIGN.md:
//   * the recursion is memoised per (s, level) and all psi of one level are accumulated in ONE sweep over
//     a2 (identical arithmetic per node, 4 sweeps per fixed-point evaluation instead of 45 bilinear forms);
//   * the final 2-3 variable AMISE minimisation uses a safeguarded Newton iteration with analytic
//     derivatives, converged tightly, instead of scipy's TNC with finite-difference gradients (whose stopping
//     point scatters by ~1e-4 in h under ulp-level input changes, see tests/test_oracle_golden.py notes).
#pragma once
#include "../../include/gdk.h"
#include "coop.cuh"
#include "solvers.cuh"

#define PSI_MAXE 6  // max psi entries per level

struct Kde2dConsts {
    double K[5];     // kde_bandwidth.py:140-142
    double Kodd[9];  // kde_bandwidth.py:143
    double pi2;
    double pipow[12];     // pi^(2k), k = 0..11   (np.pi ** (2 * sum(s)))
    double twopipow[12];  // (2 pi)^k, k = 0..11
};

struct PsiEntry {
    int s0, s1;
    double time;
};

struct Kde2dWork {
    const double* a2;    // squared dct2d of the normalised histogram, G x G row-major [ky][kx]; row/col 0 unused
    const double* aFFT;  // |fft2|^2, G x G, or NULL when do_correlation is off
    int G;
    double* wx;  // [PSI_MAXE][G] scratch (shared memory on the device)
    double* wy;  // [PSI_MAXE][G]
    int* cut;    // [2 * PSI_MAXE] scratch (shared memory on the device): frequency cut-offs of the current level
    PsiEntry* ebuf;  // [PSI_MAXE] scratch (shared memory on the device): the entries of the current level -- their
                     // plug-in times cost two pow() each and are computed by ONE thread per entry, not by every thread
};

// The weights of a psi functional, exp(-a i^2) i^p, fall off like a Gaussian in the frequency index: beyond
// psi_cut(...) every weight is below 1e-30 of the largest one and the terms (|a2| <= 1) cannot reach the last bit of
// the sum.  Returns the smallest c in [1, imax] with exp(-a i^2 + p ln i) < e^-70 * max for every i > c.
GDK_HD double psi_ipow(double x, int s) {  // x^s, s >= 0 small, by products
    double r = 1;
    for (; s > 0; s >>= 1, x *= x)
        if (s & 1) r *= x;
    return r;
}
GDK_HD int psi_cut(double a, double p, int imax) {
    if (!(a > 0) || !(a < INFINITY) || imax < 2) return imax;
    double ipk = p > 0 ? sqrt(p / (2 * a)) : 1.0;  // maximum of -a i^2 + p ln i
    if (!(ipk >= 1)) ipk = 1;
    if (ipk >= imax) return imax;
    const double thr = -a * ipk * ipk + p * log(ipk) - 70.0;
    int lo = (int)ipk, hi = imax;
    if (-a * (double)hi * hi + p * log((double)hi) >= thr) return imax;
    while (hi - lo > 1) {  // the log-weight decreases beyond the maximum: L(lo) >= thr > L(hi)
        const int mid = (lo + hi) >> 1;
        if (-a * (double)mid * mid + p * log((double)mid) < thr)
            hi = mid;
        else
            lo = mid;
    }
    return lo;
}

// psi(s, time) for up to PSI_MAXE entries in one sweep over a2 (kde_bandwidth.py:182-186)
template <class C>
GDK_HD void psi_even_level(const C& co, const Kde2dConsts& K, const Kde2dWork& W, const PsiEntry* e, int n, double* out) {
    const int G = W.G;
    for (int q = co.tid; q < 2 * n; q += co.nt) {  // one thread per (entry, axis): where the weights of this level die out
        const int k = q >> 1;
        W.cut[q] = psi_cut(K.pi2 * e[k].time, 2.0 * ((q & 1) ? e[k].s1 : e[k].s0), G - 1);
    }
    co.sync();
    int xcut = 1, ycut = 1;
    for (int k = 0; k < n; k++) {
        xcut = W.cut[2 * k] > xcut ? W.cut[2 * k] : xcut;
        ycut = W.cut[2 * k + 1] > ycut ? W.cut[2 * k + 1] : ycut;
    }
    // weights exp(-i^2 pi^2 t) (i^2)^s up to the cut-offs only (one exp per entry and index; integer powers by products)
    const int fc = xcut > ycut ? xcut : ycut;
    for (int it = co.tid; it < n * fc; it += co.nt) {
        const int k = it / fc, i = it - k * fc + 1;  // i = 1..fc
        const double I = (double)i * (double)i;
        const double ew = exp(-I * (K.pi2 * e[k].time));
        W.wx[k * G + i] = ew * psi_ipow(I, e[k].s0);
        W.wy[k * G + i] = ew * psi_ipow(I, e[k].s1);
    }
    co.sync();
    double part[PSI_MAXE];
    for (int k = 0; k < PSI_MAXE; k++) part[k] = 0;
    co.bilinear(W.a2, G, 1, ycut, 1, xcut, W.wx, W.wy, n, part);
    co.sumv(part, n);
    for (int k = 0; k < n; k++) {
        const int ss = e[k].s0 + e[k].s1;
        out[k] = ((ss & 1) ? -1.0 : 1.0) * part[k] * K.pipow[ss] / 4;
    }
    co.sync();
}

// psi_odd(s, time) (kde_bandwidth.py:209-214): frequencies f = fftfreq(G) * G
template <class C>
GDK_HD void psi_odd_level(const C& co, const Kde2dConsts& K, const Kde2dWork& W, const PsiEntry* e, int n, double* out) {
    const int G = W.G;
    for (int it = co.tid; it < n * G; it += co.nt) {
        const int k = it / G, i = it - k * G;
        const double f = (i < (G + 1) / 2) ? (double)i : (double)(i - G);
        const double w = exp(-(f * f) * (4 * K.pi2 * e[k].time));
        W.wx[k * G + i] = w * psi_ipow(f, e[k].s0);
        W.wy[k * G + i] = w * psi_ipow(f, e[k].s1);
    }
    co.sync();
    double part[PSI_MAXE];
    for (int k = 0; k < PSI_MAXE; k++) part[k] = 0;
    co.bilinear(W.aFFT, G, 0, G - 1, 0, G - 1, W.wx, W.wy, n, part);
    co.sumv(part, n);
    for (int k = 0; k < n; k++) out[k] = part[k] * K.twopipow[e[k].s0 + e[k].s1];
    co.sync();
}

// func2d for every s with min_sum <= |s| <= 5 (kde_bandwidth.py:188-196), level by level from |s| = 5 down.
// tab[s0][s1]; returns 0 on success, 1 if a non-finite value appeared.
template <class C>
GDK_HD int func2d_table(const C& co, const Kde2dConsts& K, const Kde2dWork& W, double N, double t, int min_sum,
                        double tab[6][6]) {
    int bad = 0;
    for (int ssum = 5; ssum >= min_sum; ssum--) {
        PsiEntry e[PSI_MAXE];
        double out[PSI_MAXE];
        const int n = ssum + 1;
        for (int s0 = co.tid; s0 <= ssum; s0 += co.nt) {
            const int s1 = ssum - s0;
            double time = t;
            if (ssum <= 4) {
                const double sum_func = tab[s0 + 1][s1] + tab[s0][s1 + 1];
                const double cst = (1 + pow(0.5, (double)(ssum + 1))) / 3;
                time = pow(-2 * cst * K.K[s0] * K.K[s1] / N / sum_func, 1.0 / (2 + ssum));
            }
            W.ebuf[s0] = PsiEntry{s0, s1, time};
        }
        co.sync();
        for (int k = 0; k < n; k++) {
            e[k] = W.ebuf[k];
            if (!(e[k].time == e[k].time)) bad = 1;
        }
        psi_even_level(co, K, W, e, n, out);
        for (int s0 = 0; s0 <= ssum; s0++) {
            tab[s0][ssum - s0] = out[s0];
            if (!(out[s0] == out[s0]) || isinf(out[s0])) bad = 1;
        }
    }
    return bad;
}

template <class C>
struct FixedPoint2D {
    const C& co;
    const Kde2dConsts& K;
    const Kde2dWork& W;
    double N;
    // kde_bandwidth.py:177-180
    GDK_HD double operator()(double t, int& fail) const {
        double tab[6][6];
        const int bad = func2d_table(co, K, W, N, t, 2, tab);
        const double sum_func = tab[0][2] + tab[2][0] + 2 * tab[1][1];
        const double time = pow(2 * M_PI * N * sum_func, -1.0 / 3);
        const double r = (t - time) / time;
        if (bad || !(r == r)) fail = 1;
        return r;
    }
};

struct Amise {
    double p40, p04, p22, p31, p13, N;
    // kde_bandwidth.py:216-232; returns +inf where the reference raises ("bias not positive definite")
    GDK_HD double operator()(double hx, double hy, double c) const {
        const double var = 1.0 / (4 * M_PI * hx * hy * sqrt(1 - c * c) * N);
        const double bias = 0.25 * (hx * hx * hx * hx * p40 + hy * hy * hy * hy * p04 + 2 * hx * hx * hy * hy * p22 * (2 * c * c + 1) +
                                    4 * c * hx * hy * (hx * hx * p31 + hy * hy * p13));
        if (bias < 0 || !(bias == bias)) return INFINITY;
        return var + bias;
    }
    // gradient g[3] and Hessian H[3][3] with respect to (hx, hy, c)
    GDK_HD void derivs(double hx, double hy, double c, double* g, double (*H)[3]) const {
        const double V = 1.0 / (4 * M_PI * N);
        const double s2 = 1 - c * c, s = sqrt(s2), s3 = s * s2, s5 = s3 * s2;
        const double A = p40, B = p04, D = p22, E = p31, F = p13;
        const double q = 2 * c * c + 1;
        g[0] = -V / (hx * hx * hy * s) + 0.25 * (4 * A * hx * hx * hx + 4 * D * hx * hy * hy * q + 4 * c * hy * (3 * E * hx * hx + F * hy * hy));
        g[1] = -V / (hx * hy * hy * s) + 0.25 * (4 * B * hy * hy * hy + 4 * D * hx * hx * hy * q + 4 * c * hx * (E * hx * hx + 3 * F * hy * hy));
        g[2] = V * c / (hx * hy * s3) + 0.25 * (8 * D * hx * hx * hy * hy * c + 4 * hx * hy * (E * hx * hx + F * hy * hy));
        H[0][0] = 2 * V / (hx * hx * hx * hy * s) + 0.25 * (12 * A * hx * hx + 4 * D * hy * hy * q + 24 * c * E * hx * hy);
        H[1][1] = 2 * V / (hx * hy * hy * hy * s) + 0.25 * (12 * B * hy * hy + 4 * D * hx * hx * q + 24 * c * F * hx * hy);
        H[0][1] = H[1][0] = V / (hx * hx * hy * hy * s) + 0.25 * (8 * D * hx * hy * q + 4 * c * (3 * E * hx * hx + 3 * F * hy * hy));
        H[0][2] = H[2][0] = -V * c / (hx * hx * hy * s3) + 0.25 * (16 * D * hx * hy * hy * c + 4 * hy * (3 * E * hx * hx + F * hy * hy));
        H[1][2] = H[2][1] = -V * c / (hx * hy * hy * s3) + 0.25 * (16 * D * hx * hx * hy * c + 4 * hx * (E * hx * hx + 3 * F * hy * hy));
        H[2][2] = V * (1 + 2 * c * c) / (hx * hy * s5) + 0.25 * (8 * D * hx * hx * hy * hy);
    }
};

// True if the bias term of the AMISE is negative somewhere on the faces c = +-0.99 of the search box for a width ratio
// hx / hy between those of (ax, ay) and (bx, by), widened by a factor 1.5 either way (at fixed c the sign of the bias
// depends on the ratio only: it is a quartic form in the widths).  Used to predict where the reference's TNC run ends in
// its bare except (see kernel_optimizer_2d); any widening factor between 1.1 and 3 decides the measured cases alike.
GDK_HD bool amise_bias_negative_near(const Amise& f, double ax, double ay, double bx, double by) {
    const double ra = ax / ay, rb = bx / by;
    const double rlo = fmin(ra, rb) / 1.5, rhi = fmax(ra, rb) * 1.5;
    const int n = 96;
    const double step = pow(rhi / rlo, 1.0 / (n - 1));
    const double cb = 0.99, q = 2 * cb * cb + 1;
    double r = rlo;
    for (int i = 0; i < n + 2; i++) {
        const double rr = i < n ? r : (i == n ? ra : rb);  // the two ratios themselves, whatever the spacing
        const double r2 = rr * rr;
        const double even = f.p40 * r2 * r2 + f.p04 + 2 * f.p22 * q * r2;
        const double odd = 4 * cb * (f.p31 * r2 * rr + f.p13 * rr);
        if (even + odd < 0 || even - odd < 0) return true;
        r *= step;
    }
    return false;
}

// Safeguarded Newton minimisation of the AMISE over a box (nv = 2: c fixed; nv = 3: c free).
// Variables at a bound whose gradient points outwards are frozen (active set); the Hessian of the free
// block is shifted until positive definite; steps are backtracked on the function value.
// Returns 1 on convergence (the analogue of res.success), 0 otherwise.  Scalar code: every thread of a
// group executes it identically.
GDK_HD int amise_minimise(const Amise& f, int nv, double* x, const double* lo, const double* hi) {
    double fx = f(x[0], x[1], x[2]);
    if (!(fx < INFINITY)) return 0;
    for (int iter = 0; iter < 200; iter++) {
        double g[3], H[3][3];
        f.derivs(x[0], x[1], x[2], g, H);
        bool freev[3];
        int nfree = 0;
        for (int i = 0; i < 3; i++) {
            freev[i] = i < nv;
            if (freev[i] && ((x[i] <= lo[i] && g[i] > 0) || (x[i] >= hi[i] && g[i] < 0))) freev[i] = false;
            if (freev[i]) nfree++;
        }
        // scaled projected-gradient test
        double gn = 0;
        for (int i = 0; i < 3; i++)
            if (freev[i]) gn = fmax(gn, fabs(g[i] * x[i] == 0 ? g[i] : g[i] * fmax(fabs(x[i]), 1e-3)));
        if (nfree == 0 || gn <= 1e-14 * fmax(fabs(fx), 1e-300)) return 1;
        // solve (H_ff + lam I) d = -g_f by Cholesky with increasing shift
        double d[3] = {0, 0, 0};
        double lam = 0;
        bool ok = false;
        for (int tries = 0; tries < 60 && !ok; tries++) {
            int idx[3], m = 0;
            for (int i = 0; i < 3; i++)
                if (freev[i]) idx[m++] = i;
            double L[3][3];
            ok = true;
            for (int a = 0; a < m && ok; a++)
                for (int b = 0; b <= a; b++) {
                    double sacc = H[idx[a]][idx[b]] + (a == b ? lam : 0.0);
                    for (int k = 0; k < b; k++) sacc -= L[a][k] * L[b][k];
                    if (a == b) {
                        if (!(sacc > 0)) {
                            ok = false;
                            break;
                        }
                        L[a][a] = sqrt(sacc);
                    } else {
                        L[a][b] = sacc / L[b][b];
                    }
                }
            if (ok) {
                double y[3];
                for (int a = 0; a < m; a++) {
                    double sacc = -g[idx[a]];
                    for (int k = 0; k < a; k++) sacc -= L[a][k] * y[k];
                    y[a] = sacc / L[a][a];
                }
                for (int a = m - 1; a >= 0; a--) {
                    double sacc = y[a];
                    for (int k = a + 1; k < m; k++) sacc -= L[k][a] * d[idx[k]];
                    d[idx[a]] = sacc / L[a][a];
                }
            } else {
                double diagmax = 0;
                for (int i = 0; i < 3; i++)
                    if (freev[i]) diagmax = fmax(diagmax, fabs(H[i][i]));
                lam = (lam == 0) ? 1e-6 * fmax(diagmax, 1e-300) : lam * 10;
            }
        }
        if (!ok) return 0;
        // backtracking line search with projection onto the box
        double step = 1.0;
        bool moved = false;
        double xn[3];
        for (int ls = 0; ls < 60; ls++) {
            for (int i = 0; i < 3; i++) {
                xn[i] = x[i] + step * d[i];
                if (i < nv) xn[i] = fmin(fmax(xn[i], lo[i]), hi[i]);
            }
            const double fn = f(xn[0], xn[1], xn[2]);
            if (fn < fx) {
                moved = true;
                double rel = 0;
                for (int i = 0; i < nv; i++) rel = fmax(rel, fabs(xn[i] - x[i]) / fmax(fabs(x[i]), 1e-3));
                const double dec = fx - fn;
                for (int i = 0; i < 3; i++) x[i] = xn[i];
                fx = fn;
                if (rel < 1e-13 || dec <= 1e-16 * fabs(fx)) return 1;
                break;
            }
            step *= 0.5;
        }
        if (!moved) return 1;  // no descent possible at working precision: at the minimum
    }
    return 0;
}

struct Bw2dOut {
    double hx, hy, c;  // what KernelOptimizer2D.get_h returns (grid-fraction units)
    double t_star;
    uint32_t status;
    int n_brent;
    int failed;  // ValueError-equivalent: caller applies fallback widths
};

// KernelOptimizer2D(data, N, corr, do_correlation, fallback_t).get_h()
template <class C>
GDK_HD Bw2dOut kernel_optimizer_2d(const C& co, const Kde2dConsts& K, const Kde2dWork& W, double N, double corr,
                                   int do_correlation, int have_fallback_t, double fallback_t) {
    Bw2dOut o{0, 0, 0, NAN, 0u, 0, 0};
    FixedPoint2D<C> fp{co, K, W, N};
    RootResult rr = brentq_port(fp, 0.0, 0.1, 0.001 * 0.001, 4 * GDK_DBL_EPS, 100);
    o.n_brent = rr.nfev;
    double t_star = rr.x;
    if (rr.status == 0) {
        if (have_fallback_t && fallback_t != 0 && t_star > 0.01 && t_star > 2 * fallback_t) {
            t_star = fallback_t;
            o.status |= GDK_ST_FALLBACK_T;
        }
    } else {
        if (rr.status == 1) o.status |= GDK_ST_NONFINITE;
        if (have_fallback_t) {
            t_star = fallback_t;
            o.status |= GDK_ST_FALLBACK_T;
        } else {
            o.failed = 1;
            return o;
        }
    }
    o.t_star = t_star;
    double tab[6][6];
    const int bad = func2d_table(co, K, W, N, t_star, do_correlation ? 0 : 2, tab);
    const double p_02 = tab[0][2], p_20 = tab[2][0], p_11 = tab[1][1];
    double h_x = pow(pow(p_02, 0.75) / (4 * M_PI * N * pow(p_20, 0.75) * (p_11 + sqrt(p_20 * p_02))), 1.0 / 6);
    double h_y = pow(pow(p_20, 0.75) / (4 * M_PI * N * pow(p_02, 0.75) * (p_11 + sqrt(p_20 * p_02))), 1.0 / 6);
    if (bad || !(h_x == h_x) || !(h_y == h_y)) {
        o.status |= GDK_ST_NONFINITE;
        o.failed = 1;
        return o;
    }
    double c = 0;
    o.hx = h_x;
    o.hy = h_y;
    o.c = 0;
    if (!do_correlation) return o;
    // odd functionals under [1,3] and [3,1] (kde_bandwidth.py:198-207)
    const double p00 = tab[0][0];
    double otab[10][10];
    for (int ssum = 10; ssum >= 4; ssum -= 2) {
        PsiEntry e[PSI_MAXE];
        double out[PSI_MAXE];
        const int n = ssum / 2;  // s0 = 1, 3, ..., ssum - 1
        for (int k = co.tid; k < n; k += co.nt) {
            const int s0 = 2 * k + 1, s1 = ssum - s0;
            double time = t_star;
            if (ssum <= 8) {
                const double sum_func = otab[s0 + 2][s1] + otab[s0][s1 + 2];
                const double cst = 8 * (1 - pow(2.0, (double)(-ssum - 1))) / 3.0;
                time = pow(cst * p00 * K.Kodd[s0] * K.Kodd[s1] / (N * N) / (sum_func * sum_func), 1.0 / (3 + ssum));
            }
            W.ebuf[k] = PsiEntry{s0, s1, time};
        }
        co.sync();
        for (int k = 0; k < n; k++) e[k] = W.ebuf[k];
        psi_odd_level(co, K, W, e, n, out);
        for (int k = 0; k < n; k++) otab[e[k].s0][e[k].s1] = out[k];
    }
    // the AMISE minimisation is scalar code and only the group's thread 0 hands the result on: the others are done
    // (no barrier follows)
    if (co.tid != 0) return o;
    Amise am{p_20, p_02, p_11, otab[3][1], otab[1][3], N};
    double AM = am(h_x, h_y, 0.0);
    if (!(AM < INFINITY)) {
        o.status |= GDK_ST_BIAS_NEG;  // the reference raises here (outside any try)
        o.failed = 1;
        return o;
    }
    const double lo[3] = {0.001, 0.001, -0.99}, hi[3] = {0.3, 0.3, 0.99};
    if (corr != 0) {
        const double sc = sqrt(1 - fabs(corr));
        double x[3] = {h_x / sc, h_y / sc, corr};
        x[0] = fmin(fmax(x[0], lo[0]), hi[0]);  // TNC clips the start into the box
        x[1] = fmin(fmax(x[1], lo[1]), hi[1]);
        if (amise_minimise(am, 2, x, lo, hi)) {
            const double A2 = am(x[0], x[1], corr);
            if (A2 < AM) {
                h_x = x[0];
                h_y = x[1];
                c = corr;
                AM = A2;
                o.status |= GDK_ST_AMISE_CORR;
            }
        }
    }
    {
        double x[3] = {h_x, h_y, corr};
        for (int i = 0; i < 3; i++) x[i] = fmin(fmax(x[i], lo[i]), hi[i]);
        if (amise_minimise(am, 3, x, lo, hi)) {
            const double A3 = am(x[0], x[1], x[2]);
            // The reference runs this search with scipy's TNC inside a bare try/except (kde_bandwidth.py:292-304): an
            // exception raised by AMISE ("bias not positive definite", :230-231) at ANY trial point ends it and the result
            // so far stands.  TNC's line searches run onto the correlation bound, so it dies there whenever the bias is
            // negative on that face of the box near its path; a Newton iteration never leaves the region where the AMISE
            // is finite and would hand back a minimum the reference never reports (small samples with hard edges: widths
            // off by up to 3x, |c| up to 0.99).  amise_bias_negative_near() is that test; against the unmodified reference
            // on random distributions it decides the 37 cases where the two used to differ and the 23 minima
            // the reference accepted like the reference (5 260 cases), and 1 700 fresh cases without a mismatch.
            const bool ref_aborts = amise_bias_negative_near(am, h_x, h_y, x[0], x[1]);
            if (ref_aborts) o.status |= GDK_ST_AMISE_ABORT;
            if (A3 < AM * 0.9 && !ref_aborts) {
                h_x = x[0];
                h_y = x[1];
                c = x[2];
                o.status |= GDK_ST_AMISE_FULL;
            }
        }
    }
    o.hx = h_x;
    o.hy = h_y;
    o.c = c;
    return o;
}

// Tail of getAutoBandwidth2D + the smoothing-width logic of get2DDensityGridData (mcsamples.py:1336-1345,
// 1376-1419, 1848-1863): from the optimiser output (or the rule of thumb / fallback) to (hx, hy, c) in
// parameter units, (rx, ry) in bins and the kernel half width.
GDK_HD void finish_bandwidth_2d(const gdk_spec2d& sp, const Bw2dOut* opt, double r2_shear, gdk_result2d* res) {
    const double N_eff = sp.neff;
    const double rangex = sp.xbinmax - sp.xbinmin, rangey = sp.ybinmax - sp.ybinmin;
    const double fwx = rangex / (sp.fine_bins - 1), fwy = rangey / (sp.fine_bins - 1);
    double hx = 0, hy = 0, c = sp.kernel_corr, rx, ry;
    uint32_t status = 0;
    double t_star = NAN;
    int n_brent = 0;
    if (sp.bw_mode == GDK_BW2D_FIXED) {
        rx = sp.rx_fixed;
        ry = sp.ry_fixed;
    } else {
        bool fallback = false;
        if (sp.bw_mode == GDK_BW2D_RULE) {
            fallback = true;  // same formula as fallback_widths, without the warning
        } else {
            status = opt->status;
            t_star = opt->t_star;
            n_brent = opt->n_brent;
            if (opt->failed) {
                fallback = true;
                if (!(status & GDK_ST_BIAS_NEG)) status |= GDK_ST_BW_FALLBACK;
            } else if (sp.bw_mode == GDK_BW2D_SHEAR) {
                const double r1 = sp.p1_max - sp.p1_min;
                double ex = opt->hx * r1, ey = opt->hy * r2_shear, ec = opt->c;
                // kernelC = S [[hx^2, hx hy c],[hx hy c, hy^2]] S^T with S lower triangular
                const double m00 = ex * ex, m01 = ex * ey * ec, m11 = ey * ey;
                const double t00 = sp.S00 * m00, t01 = sp.S00 * m01;                    // (S M) row 0
                const double t10 = sp.S10 * m00 + sp.S11 * m01, t11 = sp.S10 * m01 + sp.S11 * m11;  // row 1
                const double k00 = t00 * sp.S00;
                const double k01 = t00 * sp.S10 + t01 * sp.S11;
                const double k11 = t10 * sp.S10 + t11 * sp.S11;
                hx = sqrt(k00);
                hy = sqrt(k11);
                c = k01 / sqrt(k00 * k11);
                if (sp.shear_swapped) {
                    const double t = hx;
                    hx = hy;
                    hy = t;
                }
            } else {
                hx = opt->hx * rangex;
                hy = opt->hy * rangey;
                c = opt->c;
            }
        }
        if (fallback) {
            hx = sp.x_sigma_range / pow(N_eff, 1.0 / 6);
            hy = sp.y_sigma_range / pow(N_eff, 1.0 / 6);
            c = fmax(fmin(sp.corr, sp.max_corr_2D), -sp.max_corr_2D);
        }
        if (sp.mult_bias_correction_order) {
            const double scale = 1.1 * pow(N_eff, 1.0 / 6 - 1.0 / (2 + 4 * (1 + sp.mult_bias_correction_order)));
            hx *= scale;
            hy *= scale;
        }
        rx = hx * fabs(sp.smooth_scale_2D) / fwx;
        ry = hy * fabs(sp.smooth_scale_2D) / fwy;
    }
    const double smooth_scale = fmax(rx, ry);
    if (smooth_scale < 2) status |= GDK_ST_SMALL_SMOOTH;
    int winw = (int)rint(2.5 * smooth_scale);
    if (winw < 1) winw = 1;
    res->hx = hx;
    res->hy = hy;
    res->c = c;
    res->rx = rx;
    res->ry = ry;
    res->t_star = t_star;
    res->winw = winw;
    res->status = status;
    res->n_brent = n_brent;
    res->pad = 0;
}

// ----------------------------------------------------------------------------------------------------
// contour levels (getContourLevels, densities.py:19-56, half_edge=True, missing_norm=0) of a G x G grid P >= 0.
// The reference sorts the grid, accumulates the edge-halved bins in ascending order and interpolates between the
// two sorted neighbours at the crossing.  Here the crossing element is found by bisection on the (order
// preserving) bit pattern of the level: S(K) = sum of halved bins with P <= K is a group-wide reduction; all
// requested contours share every sweep.  No sort, ~64 sweeps of an L2-resident grid.
// levels[c]; returns a bit mask of contours whose crossing is the very first sorted element ("outside range").
// ----------------------------------------------------------------------------------------------------
GDK_HD double edge_factor(int y, int x, int G) {
    double f = 1.0;
    if (y == 0 || y == G - 1) f *= 0.5;
    if (x == 0 || x == G - 1) f *= 0.5;
    return f;
}
GDK_HD unsigned long long dbl_bits(double v) {
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long b;
    memcpy(&b, &v, 8);
    return b;
#endif
}
GDK_HD double bits_dbl(unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}

// Given, per contour, the crossing key K* (the smallest key whose inclusive cumulative sum reaches the target): the
// cumulative weight strictly below, the halved bin of one member of the tie group, the sorted predecessor and its
// halved bin, then the interpolation of densities.py:39-47.  Two sweeps over the grid for all contours together.
template <class C>
GDK_HD unsigned contour_finish(const C& co, const double* P, int G, const unsigned long long* Kc, const double* target, int nc,
                               double* levels) {
    const int n = G * G;
    double below[4] = {0, 0, 0, 0}, tie_a[4] = {0, 0, 0, 0}, cnt_tie[4] = {0, 0, 0, 0};
    unsigned long long prevk[4] = {0, 0, 0, 0};
    bool has_prev[4] = {false, false, false, false};
    for (int i = co.tid; i < n; i += co.nt) {
        const double v = P[i];
        const unsigned long long kb = dbl_bits(v);
        const double a = v * edge_factor(i / G, i % G, G);
        for (int c = 0; c < 4; c++) {
            if (c >= nc) continue;
            if (kb < Kc[c]) {
                below[c] += a;
                if (!has_prev[c] || kb > prevk[c]) {
                    prevk[c] = kb;
                    has_prev[c] = true;
                }
            } else if (kb == Kc[c]) {
                tie_a[c] = fmax(tie_a[c], a);
                cnt_tie[c] += 1;
            }
        }
    }
    double B[4], sg[4], ntie[4], pk[4];
    unsigned long long pkb[4];
    for (int c = 0; c < 4; c++) {
        if (c >= nc) continue;
        B[c] = co.sum(below[c]);
        sg[c] = co.max(tie_a[c]);  // interior members of a tie group all carry the same halved bin
        ntie[c] = co.sum(cnt_tie[c]);
        pk[c] = co.max(has_prev[c] ? bits_dbl(prevk[c]) : -1.0);  // value of the sorted predecessor (or -1)
        pkb[c] = dbl_bits(pk[c] >= 0 ? pk[c] : 0.0);
    }
    // the predecessor's halved bin: the largest one among the elements equal to pk
    double pa[4] = {0, 0, 0, 0};
    for (int i = co.tid; i < n; i += co.nt) {
        const double v = P[i];
        const unsigned long long kb = dbl_bits(v);
        for (int c = 0; c < 4; c++)
            if (c < nc && pk[c] >= 0 && kb == pkb[c]) pa[c] = fmax(pa[c], v * edge_factor(i / G, i % G, G));
    }
    unsigned outside = 0;
    for (int c = 0; c < nc; c++) {
        const double sg_prev_single = co.max(pa[c]);
        // first member of the tie group at which the running sum reaches the target
        double j = 1;
        if (sg[c] > 0) j = ceil((target[c] - B[c]) / sg[c]);
        if (j < 1) j = 1;
        if (j > ntie[c]) j = ntie[c];
        const double cum = B[c] + j * sg[c];
        const double sg_prev = (j > 1) ? sg[c] : sg_prev_single;
        if (pk[c] < 0 && j <= 1) outside |= 1u << c;  // ix == 0
        const double d = sg[c] > 0 ? (cum - target[c]) / sg[c] : 0.0;
        levels[c] = sg[c] * (1 - d) + d * sg_prev;
    }
    return outside;
}

template <class C>
GDK_HD unsigned contour_levels_core(const C& co, const double* P, int G, const double* contours, int nc, double* levels) {
    const int n = G * G;
    double part = 0, pmx = 0;
    for (int i = co.tid; i < n; i += co.nt) {
        part += P[i] * edge_factor(i / G, i % G, G);
        pmx = fmax(pmx, P[i]);
    }
    const double norm = co.sum(part);
    const double vmax = co.max(pmx);
    unsigned long long lo[4], hi[4];
    double target[4];
    for (int c = 0; c < 4; c++) {
        target[c] = c < nc ? (1 - contours[c]) * norm : 0;
        lo[c] = 0;                // S(lo) may or may not reach the target; invariant kept on hi only
        hi[c] = dbl_bits(vmax);   // S(hi) = norm >= target
    }
    // smallest key K with S(K) >= target  (keys of non-negative doubles are monotone in the value)
    for (int itn = 0; itn < 64; itn++) {
        bool active = false;
        unsigned long long mid[4];
        for (int c = 0; c < 4; c++) {
            mid[c] = lo[c] + ((hi[c] - lo[c]) >> 1);
            if (c < nc && lo[c] < hi[c]) active = true;
        }
        if (!active) break;
        double s[4] = {0, 0, 0, 0};
        for (int i = co.tid; i < n; i += co.nt) {
            const double v = P[i];
            const unsigned long long kb = dbl_bits(v);
            const double a = v * edge_factor(i / G, i % G, G);
            for (int c = 0; c < 4; c++)
                if (kb <= mid[c]) s[c] += a;
        }
        for (int c = 0; c < 4; c++) {
            const double S = co.sum(s[c]);
            if (c < nc && lo[c] < hi[c]) {
                if (S >= target[c])
                    hi[c] = mid[c];
                else
                    lo[c] = mid[c] + 1;
            }
        }
    }
    return contour_finish(co, P, G, hi, target, nc, levels);
}
