// kde1d_core.cuh -- grid-sized part of one 1D density, executed by one cooperating group (one CTA).
//
// Restates, for the device: gaussian_kde_bandwidth_binned + _bandwidth_fixed_point
// (kde_bandwidth.py:59-73, 102-135), getAutoBandwidth1D (mcsamples.py:1237-1283), the smoothing-width
// logic and Kernel1D (mcsamples.py:1563-1589, 129-135), convolve1D as direct convolution
// (convolve.py:196-202), the linear/quadratic boundary kernels (mcsamples.py:1600-1647), the
// multiplicative bias correction (mcsamples.py:1649-1666) and normalize('max') (densities.py:71-92).
#pragma once
#include "../../include/gdk.h"
#include "fft.cuh"
#include "solvers.cuh"

// constants of the ISJ recursion, filled on the host with the same libm expressions numpy evaluates
struct IsjConsts {
    double two_pi_pow[8];  // 2 * pi^(2j), j = 0..7
    double cj[8];          // _kde_consts_1d indexed by j (j = 2..6)
    double pi2;            // pi^2
    double rootpi;
};

struct Kde1dWork {
    // all arrays of length >= F unless noted; may live in shared or global memory
    double* bins;  // weighted histogram (input, preserved)
    double* a2;    // (dct/2)^2, index k-1 for k = 1..F-1
    double* logI;  // log(k^2)
    double* P;     // density being built
    double* aux;   // scratch: fine = bins/P ; xP ; conv results
    double* aux2;  // scratch
    double* win;   // kernel taps, length 2*winw+1 <= F
    cplx* ca;      // FFT ping
    cplx* cb;      // FFT pong
    const cplx* tw;      // n-th roots (pow2 path) or NULL
    const cplx* tw4;     // 4n-th roots, first n (pow2 path)
    const double* cos4;  // cos(2 pi j/4n) table (direct path)
    // mean likelihoods (meanlikes=True, mcsamples.py:1556-1561, 1672-1684); all NULL when not requested
    const double* likebins = nullptr;  // histogram of weights * exp(mean_loglike - loglikes)
    double* raw = nullptr;             // scratch: the first convolution ("rawbins", mcsamples.py:1597-1598)
    double* likes_out = nullptr;       // max-normalised mean likelihoods
};

template <class C>
struct IsjFixedPoint {
    const C& co;
    const IsjConsts& K;
    const double* a2;
    const double* logI;
    int F;
    double N;
    // kde_bandwidth.py:59-73
    GDK_HD double operator()(double h, int& fail) const {
        if (h <= 0) return h - 1;
        double f = dotexp(7, K.pi2 * h * h);
        f = K.two_pi_pow[7] * f;
        for (int j = 6; j >= 2; j--) {
            const double tj = pow(K.cj[j] / N / f, 2.0 / (3.0 + 2 * j));
            f = K.two_pi_pow[j] * dotexp(j, K.pi2 * tj);
            if (!(f != 0) || f != f || isinf(f)) {  // "zero f" exception (or non-finite)
                fail = 1;
                return 0;
            }
        }
        return h - pow(2 * N * K.rootpi * f, -1.0 / 5);
    }
    // sum_k a2[k] exp(j logI_k - I_k * s)
    GDK_HD double dotexp(int j, double s) const {
        double acc = 0;
        for (int k = 1 + co.tid; k < F; k += co.nt) {
            const double I = (double)k * (double)k;
            acc += a2[k - 1] * exp(j * logI[k - 1] - I * s);
        }
        return co.sum(acc);
    }
};

// 'same' convolution of a length-F signal with a centred kernel of half width w (taps win[u+w], u=-w..w):
// out[i] = sum_u win(u) * x[i-u], zero padded (np.convolve(x, win, 'same') for 2w+1 <= F).
template <class C>
GDK_HD void conv_same(const C& co, const double* x, const double* win, int w, int F, double* out) {
    for (int i = co.tid; i < F; i += co.nt) {
        const int ulo = (i - (F - 1) > -w) ? i - (F - 1) : -w;
        const int uhi = (i < w) ? i : w;
        double acc = 0;
        for (int u = ulo; u <= uhi; u++) acc += win[u + w] * x[i - u];
        out[i] = acc;
    }
    co.sync();
}

// circular convolution of convolve1D_periodic (convolve.py:326-367): the last bin is folded onto the first,
// period F-1, centred kernel; the result is extended by its first element.
template <class C>
GDK_HD void conv_circ(const C& co, const double* x, const double* win, int w, int F, double* out) {
    const int Nc = F - 1;
    for (int i = co.tid; i < F; i += co.nt) {
        const int ic = i == Nc ? 0 : i;
        double acc = 0;
        for (int u = -w; u <= w; u++) {
            int s = (ic - u) % Nc;
            if (s < 0) s += Nc;
            const double v = x[s] + (s == 0 ? x[Nc] : 0.0);
            acc += win[u + w] * v;
        }
        out[i] = acc;
    }
    co.sync();
}

template <class C>
GDK_HD void kde1d_core(const C& co, const gdk_spec1d& sp, const IsjConsts& K, Kde1dWork& W, double* P_out,
                       gdk_result1d* res) {
    const int F = sp.fine_bins;
    const double fine_width = (sp.binmax - sp.binmin) / (F - 1);
    const double paramrange = sp.range_max - sp.range_min;
    uint32_t status = 0;
    double kde_h = NAN, h_raw = NAN;
    int nfev = 0;
    double smooth_1D;

    if (sp.smooth_scale_1D <= 0) {
        // ---- ISJ bandwidth on the DCT of the normalised histogram --------------------------------
        double part = 0;
        for (int i = co.tid; i < F; i += co.nt) part += W.bins[i];
        const double total = co.sum(part);
        for (int i = co.tid; i < F; i += co.nt) W.aux[i] = W.bins[i] / total;
        co.sync();
        if (W.tw)
            dct2_lines_pow2(co, W.aux, W.aux2, W.ca, W.cb, F, 1, W.tw, W.tw4);
        else
            dct2_lines_direct(co, W.aux, W.aux2, F, 1, W.cos4);
        for (int k = 1 + co.tid; k < F; k += co.nt) {
            const double a = W.aux2[k] / 2;
            W.a2[k - 1] = a * a;
            W.logI[k - 1] = log((double)k * (double)k);
        }
        co.sync();
        const double neff = sp.neff;
        const double n_scaling = pow(neff, -1.0 / 5);
        IsjFixedPoint<C> fp{co, K, W.a2, W.logI, F, neff};
        const double h0 = 0.53 * n_scaling;
        RootResult rr = hybrd1_port(fp, h0, h0 / 20, 1.0, 400);
        nfev = rr.nfev;
        bool none = rr.status != 0;
        double h = rr.x;
        if (!none && h < 0.019 * n_scaling) {
            // second-root guard, kde_bandwidth.py:124-131 (any exception inside keeps the fsolve value)
            const double xtol = h / 20;
            if (xtol > 0) {
                RootResult rb = brentq_port(fp, 0.019 * n_scaling, 0.5, xtol, 4 * GDK_DBL_EPS, 100);
                nfev += rb.nfev;
                if (rb.status == 0) {
                    h = rb.x;
                    status |= GDK_ST_USED_BRENT;
                }
            }
        }
        if (none) status |= GDK_ST_BW_FAILED_NONE;
        h_raw = none ? NAN : h;
        // getAutoBandwidth1D, mcsamples.py:1257-1283
        const double bin_range = fmax(sp.param_max, sp.range_max) - fmin(sp.param_min, sp.range_min);
        if (none || h < 0.01 * n_scaling * (sp.range_max - sp.range_min) / bin_range) {
            h = 1.06 * sp.sigma_range * n_scaling / bin_range;
            status |= GDK_ST_BW_FALLBACK;
        }
        kde_h = h;
        int m = sp.mult_bias_correction_order;
        if (sp.boundary_correction_order > 1) m = m > 1 ? m : 1;
        if (m) h = h * pow(neff, 1.0 / 5 - 1.0 / (4 * m + 5));
        double bandwidth = h * (sp.binmax - sp.binmin);
        bandwidth = fmin(bandwidth, paramrange / 4);
        smooth_1D = bandwidth * fabs(sp.smooth_scale_1D) / fine_width;
    } else if (sp.smooth_scale_1D < 1.0) {
        smooth_1D = sp.smooth_scale_1D * sp.err / fine_width;
    } else {
        smooth_1D = sp.smooth_scale_1D * sp.width / fine_width;
    }
    if (smooth_1D < 2) status |= GDK_ST_SMALL_SMOOTH;
    smooth_1D = fmin(fmax(1.0, smooth_1D), (double)(F / 2));
    const bool periodic = sp.periodic != 0;
    int winw = (int)rint(2.5 * smooth_1D);  // Python round(): half to even
    const int wcap = (periodic ? F - 1 : F) / 2 - 2;
    if (winw > wcap) winw = wcap;
    const int w = winw;

    // ---- Kernel1D, mcsamples.py:129-135 ----------------------------------------------------------
    double part = 0;
    for (int u = -w + co.tid; u <= w; u += co.nt) {
        const double t = (double)u / smooth_1D;
        const double v = exp(-(t * t) / 2.0);
        W.win[u + w] = v;
        part += v;
    }
    const double wsum = co.sum(part);
    for (int u = co.tid; u <= 2 * w; u += co.nt) W.win[u] = W.win[u] / wsum;
    co.sync();

    const bool bot = sp.has_limits_bot != 0, top = sp.has_limits_top != 0;
    const int bco = sp.boundary_correction_order;
    // ---- P = bins (*) Win, plus the boundary-kernel moments in the same sweep ----------------------
    if (periodic) {
        conv_circ(co, W.bins, W.win, w, F, W.P);
        if (W.raw) {
            for (int i = co.tid; i < F; i += co.nt) W.raw[i] = W.P[i];
            co.sync();
        }
    } else if ((bot || top) && bco >= 0) {
        // prior mask of length F+2w: 0 outside the bounded side, 1/2 on the boundary bin, 1 inside.
        // 'valid' conv: a_r[i] = sum_u u^r win(u) mask[i + w - u];   'same' conv: xP, x2P on bins.
        for (int i = co.tid; i < F; i += co.nt) {
            double P = 0, xP = 0, x2P = 0, a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
            for (int u = -w; u <= w; u++) {
                const double k = W.win[u + w];
                const int s = i - u;  // index into the unpadded grid
                if (s >= 0 && s < F) {
                    const double b = W.bins[s];
                    P += k * b;
                    xP += (k * u) * b;
                    x2P += ((k * u) * u) * b;
                }
                double mk = 1.0;
                if (bot) mk = (s < 0) ? 0.0 : (s == 0 ? 0.5 : mk);
                if (top) mk = (s > F - 1) ? 0.0 : (s == F - 1 ? 0.5 * mk : mk);
                // note: a bin can be both the first and the last only for F == 1 (excluded)
                const double ku = k * u;
                a0 += k * mk;
                a1 += ku * mk;
                a2 += (ku * u) * mk;
                a3 += (ku * (double)(u * u)) * mk;          // xWin * x**2
                a4 += (ku * (double)(u * u * u)) * mk;      // xWin * x**3
            }
            double out = P;
            if (a0 * P != 0) {
                const double normed = P / a0;
                if (bco == 0) {
                    out = normed;
                } else {
                    double corrected;
                    if (bco == 1) {
                        corrected = (P * a2 - xP * a1) / (a0 * a2 - a1 * a1);
                    } else {
                        const double denom = a4 * a2 * a0 - a4 * a1 * a1 - a2 * a2 * a2 - a3 * a3 * a0 + 2 * a1 * a2 * a3;
                        const double A = a4 * a2 - a3 * a3;
                        const double B = a2 * a3 - a4 * a1;
                        const double Cq = a3 * a1 - a2 * a2;
                        corrected = (P * A + xP * B + x2P * Cq) / denom;
                    }
                    out = normed * exp(fmin(corrected / normed, 4.0) - 1);
                }
            }
            W.P[i] = out;
            if (W.raw) W.raw[i] = P;
        }
        co.sync();
    } else {
        conv_same(co, W.bins, W.win, w, F, W.P);
        if (W.raw) {
            for (int i = co.tid; i < F; i += co.nt) W.raw[i] = W.P[i];
            co.sync();
        }
        if (bco == 2) {
            // higher-order kernel for unbounded parameters, mcsamples.py:1638-1647
            double p2 = 0, p4 = 0;
            for (int u = -w + co.tid; u <= w; u += co.nt) {
                const double k2 = W.win[u + w] * (double)(u * u);  // Win * x**2
                p2 += k2;
                p4 += k2 * (double)(u * u);
            }
            const double a2 = co.sum(p2), a4 = co.sum(p4);
            for (int i = co.tid; i < F; i += co.nt) {
                const int ulo = (i - (F - 1) > -w) ? i - (F - 1) : -w;
                const int uhi = (i < w) ? i : w;
                double x2P = 0;
                for (int u = ulo; u <= uhi; u++) x2P += (W.win[u + w] * (double)(u * u)) * W.bins[i - u];
                const double P = W.P[i];
                if (P > 0) {
                    const double corrected = (P * a4 - a2 * x2P) / (a4 - a2 * a2);
                    W.aux[i] = P * exp(fmin(corrected / P, 2.0) - 1);
                } else {
                    W.aux[i] = P;
                }
            }
            co.sync();
            for (int i = co.tid; i < F; i += co.nt) W.P[i] = W.aux[i];
            co.sync();
        }
    }

    // ---- multiplicative bias correction, mcsamples.py:1649-1666 ------------------------------------
    for (int it = 0; it < sp.mult_bias_correction_order; it++) {
        for (int i = co.tid; i < F; i += co.nt) {
            const double p = W.P[i];
            W.aux[i] = W.bins[i] / (p == 0 ? 1.0 : p);
        }
        co.sync();
        if (periodic) {  // mcsamples.py:1663-1666: P *= circular conv, no normaliser
            conv_circ(co, W.aux, W.win, w, F, W.aux2);
            for (int i = co.tid; i < F; i += co.nt) W.P[i] = W.P[i] * W.aux2[i];
            co.sync();
            continue;
        }
        for (int i = co.tid; i < F; i += co.nt) {
            const int ulo = (i - (F - 1) > -w) ? i - (F - 1) : -w;
            const int uhi = (i < w) ? i : w;
            double acc = 0, a0 = 0;
            for (int u = ulo; u <= uhi; u++) {
                const double k = W.win[u + w];
                const int s = i - u;
                acc += k * W.aux[s];
                double mk = 1.0;
                if (bot && s == 0) mk *= 0.5;
                if (top && s == F - 1) mk *= 0.5;
                a0 += k * mk;
            }
            W.aux2[i] = W.P[i] * acc / a0;
        }
        co.sync();
        for (int i = co.tid; i < F; i += co.nt) W.P[i] = W.aux2[i];
        co.sync();
    }

    // ---- normalize('max') ---------------------------------------------------------------------------
    double pm = -INFINITY;
    for (int i = co.tid; i < F; i += co.nt) pm = fmax(pm, W.P[i]);
    const double mx = co.max(pm);
    if (!(mx != 0)) status |= GDK_ST_ZERO_MAX;
    for (int i = co.tid; i < F; i += co.nt) P_out[i] = (mx != 0) ? W.P[i] / mx : W.P[i];
    if (W.likebins && W.raw && W.likes_out) {
        // ---- mean likelihoods, mcsamples.py:1672-1682 (shade_likes_is_mean_loglikes = False) ---------
        for (int i = co.tid; i < F; i += co.nt) {
            const double pf = (mx != 0) ? W.P[i] / mx : W.P[i];
            W.aux[i] = pf > 0 ? W.likebins[i] / pf : W.likebins[i];
        }
        co.sync();
        if (periodic)
            conv_circ(co, W.aux, W.win, w, F, W.aux2);
        else
            conv_same(co, W.aux, W.win, w, F, W.aux2);
        double lm = -INFINITY;
        for (int i = co.tid; i < F; i += co.nt) {
            const double pf = (mx != 0) ? W.P[i] / mx : W.P[i];
            double v = W.aux2[i];
            if (pf > 0) v *= pf / W.raw[i];
            W.aux2[i] = v;
            lm = fmax(lm, v);
        }
        const double lmx = co.max(lm);
        for (int i = co.tid; i < F; i += co.nt) W.likes_out[i] = W.aux2[i] / lmx;
    }
    if (co.tid == 0) {
        res->kde_h = kde_h;
        res->h_raw = h_raw;
        res->smooth_1D = smooth_1D;
        res->winw = winw;
        res->status = status;
        res->n_feval = nfev;
        res->pad = 0;
    }
    co.sync();
}
