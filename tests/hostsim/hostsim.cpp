// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the grid-stage device headers (getdist_b200/csrc/*.cuh) for the HOST with g++ (CoopHost:
// one thread, barriers are no-ops) so that the arithmetic the CUDA kernels execute can be checked
// against the oracle in a container without a GPU.  Never loaded by the product.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../getdist_b200/csrc/host_tables.h"
#include "../../getdist_b200/csrc/kde1d_core.cuh"
#include "../../getdist_b200/csrc/kde2d_core.cuh"

extern "C" {

// scipy.optimize.fsolve / brentq restatements on a test function f(x) = a*(x-r)^3 + b*(x-r) (+ optional kink)
struct TestFn {
    double a, b, r;
    int count;
    double xs[512];
    double operator()(double x, int& fail) {
        if (count < 512) xs[count] = x;
        count++;
        double d = x - r;
        return a * d * d * d + b * d;
    }
};

int hs_hybrd1(double a, double b, double r, double x0, double xtol, double factor, double* xout, double* xs, int* nfev) {
    TestFn f{a, b, r, 0, {0}};
    RootResult rr = hybrd1_port(f, x0, xtol, factor, 400);
    *xout = rr.x;
    *nfev = rr.nfev;
    for (int i = 0; i < f.count && i < 512; i++) xs[i] = f.xs[i];
    return rr.status;
}

int hs_brentq(double a, double b, double r, double xa, double xb, double xtol, double* xout, double* xs, int* nfev) {
    TestFn f{a, b, r, 0, {0}};
    RootResult rr = brentq_port(f, xa, xb, xtol, 4 * GDK_DBL_EPS, 100);
    *xout = rr.x;
    *nfev = rr.nfev;
    for (int i = 0; i < f.count && i < 512; i++) xs[i] = f.xs[i];
    return rr.status;
}

// forward FFT / DCT-II of nl lines of length n
int hs_fft(const double* re, const double* im, int n, int nl, double* ore, double* oim) {
    CoopHost co;
    std::vector<cplx> a((size_t)n * nl), b((size_t)n * nl), tw(n), out((size_t)n * nl);
    gdk_fill_roots(tw.data(), n, n);
    for (size_t i = 0; i < a.size(); i++) a[i] = cplx{re[i], im[i]};
    if (is_pow2(n)) {
        cplx* r = fft_lines(co, a.data(), b.data(), n, nl, tw.data());
        for (size_t i = 0; i < a.size(); i++) out[i] = r[i];
    } else {
        dft_lines_direct(co, a.data(), out.data(), n, nl, tw.data());
    }
    for (size_t i = 0; i < a.size(); i++) {
        ore[i] = out[i].x;
        oim[i] = out[i].y;
    }
    return 0;
}

int hs_dct2(const double* in, int n, int nl, double* out) {
    CoopHost co;
    if (is_pow2(n)) {
        std::vector<cplx> a((size_t)n * nl), b((size_t)n * nl), tw(n), tw4(n);
        gdk_fill_roots(tw.data(), n, n);
        gdk_fill_roots(tw4.data(), 4 * n, n);
        dct2_lines_pow2(co, in, out, a.data(), b.data(), n, nl, tw.data(), tw4.data());
    } else {
        std::vector<double> c4(4 * (size_t)n);
        gdk_fill_cos(c4.data(), 4 * n);
        dct2_lines_direct(co, in, out, n, nl, c4.data());
    }
    return 0;
}

// full 1D grid stage on a given histogram
int hs_kde1d_likes(const gdk_spec1d* sp, const double* bins, const double* likebins, double* P_out, double* likes_out,
                   gdk_result1d* res);
int hs_kde1d(const gdk_spec1d* sp, const double* bins, double* P_out, gdk_result1d* res) {
    return hs_kde1d_likes(sp, bins, nullptr, P_out, nullptr, res);
}

// likebins / likes_out may be NULL (plain density)
int hs_kde1d_likes(const gdk_spec1d* sp, const double* bins, const double* likebins, double* P_out, double* likes_out,
                   gdk_result1d* res) {
    CoopHost co;
    const int F = sp->fine_bins;
    IsjConsts K;
    gdk_fill_isj_consts(&K);
    std::vector<double> b(bins, bins + F), a2(F), logI(F), P(F), aux(F), aux2(F), win(F + 1);
    std::vector<cplx> ca(F), cb(F), tw, tw4;
    std::vector<double> c4;
    Kde1dWork W{};
    W.bins = b.data();
    W.a2 = a2.data();
    W.logI = logI.data();
    W.P = P.data();
    W.aux = aux.data();
    W.aux2 = aux2.data();
    W.win = win.data();
    W.ca = ca.data();
    W.cb = cb.data();
    if (is_pow2(F)) {
        tw.resize(F);
        tw4.resize(F);
        gdk_fill_roots(tw.data(), F, F);
        gdk_fill_roots(tw4.data(), 4 * F, F);
        W.tw = tw.data();
        W.tw4 = tw4.data();
    } else {
        c4.resize(4 * (size_t)F);
        gdk_fill_cos(c4.data(), 4 * F);
        W.cos4 = c4.data();
    }
    std::vector<double> raw(F);
    if (likebins && likes_out) {
        W.likebins = likebins;
        W.raw = raw.data();
        W.likes_out = likes_out;
    }
    kde1d_core(co, *sp, K, W, P_out, res);
    return 0;
}

// squared 2D DCT-II and |FFT2|^2 of hist/sum(hist), G x G, with the same line transforms the kernels use
int hs_xform2d(const double* hist, int G, double* a2, double* aFFT) {
    CoopHost co;
    const size_t n2 = (size_t)G * G;
    double total = 0;
    for (size_t i = 0; i < n2; i++) total += hist[i];
    std::vector<double> p(n2), t1(n2), t2(n2);
    for (size_t i = 0; i < n2; i++) p[i] = hist[i] / total;
    std::vector<cplx> a(n2), b(n2), tw(G), tw4(G), c1(n2), c2(n2);
    std::vector<double> c4(4 * (size_t)G);
    gdk_fill_roots(tw.data(), G, G);
    gdk_fill_roots(tw4.data(), 4 * G, G);
    gdk_fill_cos(c4.data(), 4 * G);
    const bool p2 = is_pow2(G);
    // dct2d = dct along axis 0 then axis 1 (convolve.py:565-566); the two commute up to rounding
    if (p2) dct2_lines_pow2(co, p.data(), t1.data(), a.data(), b.data(), G, G, tw.data(), tw4.data());
    else dct2_lines_direct(co, p.data(), t1.data(), G, G, c4.data());
    for (int y = 0; y < G; y++) for (int x = 0; x < G; x++) t2[(size_t)x * G + y] = t1[(size_t)y * G + x];
    if (p2) dct2_lines_pow2(co, t2.data(), t1.data(), a.data(), b.data(), G, G, tw.data(), tw4.data());
    else dct2_lines_direct(co, t2.data(), t1.data(), G, G, c4.data());
    for (int y = 0; y < G; y++) for (int x = 0; x < G; x++) { double v = t1[(size_t)x * G + y]; a2[(size_t)y * G + x] = v * v; }
    // fft2
    for (size_t i = 0; i < n2; i++) a[i] = cplx{p[i], 0};
    cplx* r;
    if (p2) r = fft_lines(co, a.data(), b.data(), G, G, tw.data());
    else { dft_lines_direct(co, a.data(), b.data(), G, G, tw.data()); r = b.data(); }
    for (int y = 0; y < G; y++) for (int x = 0; x < G; x++) c1[(size_t)x * G + y] = r[(size_t)y * G + x];
    if (p2) r = fft_lines(co, c1.data(), c2.data(), G, G, tw.data());
    else { dft_lines_direct(co, c1.data(), c2.data(), G, G, tw.data()); r = c2.data(); }
    for (int y = 0; y < G; y++) for (int x = 0; x < G; x++) { cplx v = r[(size_t)x * G + y]; aFFT[(size_t)y * G + x] = v.x * v.x + v.y * v.y; }
    return 0;
}

// KernelOptimizer2D(...).get_h() on precomputed a2 / aFFT
int hs_bw2d(const double* a2, const double* aFFT, int G, double N, double corr, int do_corr, int have_ft, double ft,
            double* out /*hx, hy, c, t_star*/, int* iout /*status, n_brent, failed*/) {
    CoopHost co;
    Kde2dConsts K;
    gdk_fill_kde2d_consts(&K);
    std::vector<double> wx((size_t)PSI_MAXE * G), wy((size_t)PSI_MAXE * G);
    std::vector<int> cut(2 * PSI_MAXE);
    std::vector<PsiEntry> ebuf(PSI_MAXE);
    Kde2dWork W{a2, do_corr ? aFFT : nullptr, G, wx.data(), wy.data(), cut.data(), ebuf.data()};
    Bw2dOut o = kernel_optimizer_2d(co, K, W, N, corr, do_corr, have_ft, ft);
    out[0] = o.hx; out[1] = o.hy; out[2] = o.c; out[3] = o.t_star;
    iout[0] = (int)o.status; iout[1] = o.n_brent; iout[2] = o.failed;
    return 0;
}

int hs_finish_bw2d(const gdk_spec2d* sp, const double* optv /*hx,hy,c,t_star*/, const int* opti, double r2, gdk_result2d* res) {
    Bw2dOut o{optv[0], optv[1], optv[2], optv[3], (uint32_t)opti[0], opti[1], opti[2]};
    finish_bandwidth_2d(*sp, &o, r2, res);
    return 0;
}

int hs_contours(const double* P, int G, const double* contours, int nc, double* levels) {
    CoopHost co;
    return (int)contour_levels_core(co, P, G, contours, nc, levels);
}
}

// ---- trace of the ISJ fixed-point solve (the fsolve stage of kde1d_core) on a given histogram: evaluation points and
// values of hybrd1_port, for side-by-side comparison with scipy.optimize.fsolve on the same function
extern "C" int hs_isj_trace(const double* bins, int F, double neff, double* xs, double* fs, int* nfev, double* xout) {
    CoopHost co;
    IsjConsts K;
    gdk_fill_isj_consts(&K);
    std::vector<double> aux(F), aux2(F), a2(F), logI(F);
    std::vector<cplx> ca(F), cb(F), tw, tw4;
    std::vector<double> c4;
    double total = 0;
    for (int i = 0; i < F; i++) total += bins[i];
    for (int i = 0; i < F; i++) aux[i] = bins[i] / total;
    if (is_pow2(F)) {
        tw.resize(F);
        tw4.resize(F);
        gdk_fill_roots(tw.data(), F, F);
        gdk_fill_roots(tw4.data(), 4 * F, F);
        dct2_lines_pow2(co, aux.data(), aux2.data(), ca.data(), cb.data(), F, 1, tw.data(), tw4.data());
    } else {
        c4.resize(4 * (size_t)F);
        gdk_fill_cos(c4.data(), 4 * F);
        dct2_lines_direct(co, aux.data(), aux2.data(), F, 1, c4.data());
    }
    for (int k = 1; k < F; k++) {
        const double a = aux2[k] / 2;
        a2[k - 1] = a * a;
        logI[k - 1] = log((double)k * (double)k);
    }
    IsjFixedPoint<CoopHost> fp{co, K, a2.data(), logI.data(), F, neff};
    struct Traced {
        IsjFixedPoint<CoopHost>& f;
        double* xs;
        double* fs;
        int n;
        double operator()(double h, int& fail) {
            const double v = f(h, fail);
            if (n < 512) {
                xs[n] = h;
                fs[n] = v;
            }
            n++;
            return v;
        }
    } tr{fp, xs, fs, 0};
    const double h0 = 0.53 * pow(neff, -1.0 / 5);
    RootResult rr = hybrd1_port(tr, h0, h0 / 20, 1.0, 400);
    *nfev = tr.n;
    *xout = rr.x;
    return rr.status;
}
