// hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the grid-stage device headers (getdist_b200/csrc/*.cuh) for the HOST with g++ (CoopHost:
// one thread, barriers are no-ops) so that the arithmetic the CUDA kernels execute can be checked
// against the oracle in a container without a GPU.  Never loaded by the product.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../getdist_b200/csrc/host_tables.h"
#include "../../getdist_b200/csrc/kde1d_core.cuh"

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
int hs_kde1d(const gdk_spec1d* sp, const double* bins, double* P_out, gdk_result1d* res) {
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
    kde1d_core(co, *sp, K, W, P_out, res);
    return 0;
}
}
