// race_check.cpp -- TEST INFRASTRUCTURE ONLY (never loaded by the product).
//
// The grid-stage device code (getdist_b200/csrc/{fft,kde1d_core,kde2d_core}.cuh) is written once against a "Coop"
// type (coop.cuh).  hostsim.cpp instantiates it with ONE host thread to check the arithmetic.  This file instantiates
// the very same templates with a group of REAL host threads -- CoopMT: co.sync() is a pthread barrier, the reductions
// go through a shared array -- so that ThreadSanitizer can check the BARRIER STRUCTURE the CUDA kernels rely on: a
// missing or misplaced co.sync() between a write to shared scratch and another thread's read shows up as a data
// race here, in a container without a GPU.  (The one-thread host instantiation cannot see such bugs, and on the GPU
// they may stay hidden behind warp-synchronous luck.)
//
//   g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -o race_check race_check.cpp && ./race_check
//
// Exit code 0: every case ran, multi-threaded results agree with the one-thread results (reduction order differs:
// relative tolerance), and ThreadSanitizer saw nothing (it makes the process exit with 66 otherwise).
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>

#include <functional>
#include <thread>
#include <vector>

#include "../../getdist_b200/csrc/host_tables.h"
#include "../../getdist_b200/csrc/kde1d_core.cuh"
#include "../../getdist_b200/csrc/kde2d_core.cuh"

struct BlockState {
    pthread_barrier_t bar;
    int nt;
    std::vector<double> red;  // nt doubles per value, up to 8 values
    std::vector<int> flag;
    explicit BlockState(int n) : nt(n), red((size_t)8 * n), flag(n) { pthread_barrier_init(&bar, nullptr, n); }
    ~BlockState() { pthread_barrier_destroy(&bar); }
};

// One thread of a cooperating group.  Same contract as CoopBlock (coop.cuh): sum / max / any return the same value in
// every thread, accumulated in an order that does not depend on the calling thread.
static bool g_break_barriers = false;  // self-test of the checker: barriers become no-ops -> ThreadSanitizer must object

struct CoopMT {
    int tid, nt;
    BlockState* B;
    void sync() const {
        if (!g_break_barriers) pthread_barrier_wait(&B->bar);
    }
    double sum(double v) const {
        B->red[tid] = v;
        sync();
        double t = 0;
        for (int i = 0; i < nt; i++) t += B->red[i];
        sync();  // everybody has read before the array is written again
        return t;
    }
    double max(double v) const {
        B->red[tid] = v;
        sync();
        double t = B->red[0];
        for (int i = 1; i < nt; i++) t = fmax(t, B->red[i]);
        sync();
        return t;
    }
    int any(int v) const {
        B->flag[tid] = v;
        sync();
        int t = 0;
        for (int i = 0; i < nt; i++) t |= B->flag[i];
        sync();
        return t;
    }
    void sumv(double* v, int n) const {
        for (int k = 0; k < n; k++) B->red[(size_t)k * nt + tid] = v[k];
        sync();
        for (int k = 0; k < n; k++) {
            double t = 0;
            for (int i = 0; i < nt; i++) t += B->red[(size_t)k * nt + i];
            v[k] = t;
        }
        sync();
    }
    // per-thread partial sums over this thread's share of the window: rows dealt to warps, lanes stride over x (the
    // assignment of CoopBlock::bilinear's plain path)
    void bilinear(const double* A, int G, int y0, int y1, int x0, int x1, const double* wx, const double* wy, int n,
                  double* part) const {
        const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
        for (int y = y0 + wid; y <= y1; y += nw)
            for (int x = x0 + lane; x <= x1; x += 32)
                for (int k = 0; k < n; k++) part[k] += (wy[k * G + y] * A[(size_t)y * G + x]) * wx[k * G + x];
    }
};

// run `body(co)` on nt threads forming one group
template <class F>
static void run_group(int nt, F body) {
    BlockState B(nt);
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; t++) th.emplace_back([&, t] { body(CoopMT{t, nt, &B}); });
    for (auto& t : th) t.join();
}

static int g_fail = 0;
static void expect_close(const char* what, const double* a, const double* b, size_t n, double rtol) {
    double scale = 0, worst = 0;
    for (size_t i = 0; i < n; i++) scale = fmax(scale, fabs(b[i]));
    for (size_t i = 0; i < n; i++) worst = fmax(worst, fabs(a[i] - b[i]));
    const bool ok = worst <= rtol * (scale > 0 ? scale : 1.0);
    printf("%-58s max|mt - st| = %.3e (scale %.3e) %s\n", what, worst, scale, ok ? "ok" : "MISMATCH");
    if (!ok) g_fail = 1;
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {  // xorshift64*: the inputs only need to be generic
    rng_state ^= rng_state >> 12;
    rng_state ^= rng_state << 25;
    rng_state ^= rng_state >> 27;
    return (double)((rng_state * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
}
static double nrand() { return sqrt(-2 * log(urand() + 1e-300)) * cos(2 * M_PI * urand()); }

// ---------------------------------------------------------------------------------------------------------------
// line transforms
// ---------------------------------------------------------------------------------------------------------------
static void check_transforms(int n, int nl, int nt) {
    std::vector<cplx> in((size_t)n * nl), tw(n), tw4(n);
    std::vector<double> rin((size_t)n * nl), c4(4 * (size_t)n);
    for (auto& v : in) v = cplx{nrand(), nrand()};
    for (auto& v : rin) v = nrand();
    gdk_fill_roots(tw.data(), n, n);
    gdk_fill_roots(tw4.data(), 4 * n, n);
    gdk_fill_cos(c4.data(), 4 * n);
    char name[128];
    auto flat = [](const std::vector<cplx>& v) { return reinterpret_cast<const double*>(v.data()); };
    if (is_pow2(n)) {
        std::vector<cplx> a1 = in, b1(in.size()), a2 = in, b2(in.size()), o1(in.size()), o2(in.size());
        CoopHost one;
        cplx* r1 = fft_lines(one, a1.data(), b1.data(), n, nl, tw.data());
        o1.assign(r1, r1 + in.size());
        cplx* r2 = nullptr;
        run_group(nt, [&](CoopMT co) {
            cplx* r = fft_lines(co, a2.data(), b2.data(), n, nl, tw.data());
            if (co.tid == 0) r2 = r;
        });
        o2.assign(r2, r2 + in.size());
        snprintf(name, sizeof name, "fft_lines n=%d lines=%d threads=%d", n, nl, nt);
        expect_close(name, flat(o2), flat(o1), 2 * in.size(), 1e-13);
        std::vector<double> d1(rin.size()), d2(rin.size());
        dct2_lines_pow2(one, rin.data(), d1.data(), a1.data(), b1.data(), n, nl, tw.data(), tw4.data());
        run_group(nt, [&](CoopMT co) { dct2_lines_pow2(co, rin.data(), d2.data(), a2.data(), b2.data(), n, nl, tw.data(), tw4.data()); });
        snprintf(name, sizeof name, "dct2_lines_pow2 n=%d lines=%d threads=%d", n, nl, nt);
        expect_close(name, d2.data(), d1.data(), d1.size(), 1e-13);
        // in place (the kernels transform a tile where it lies)
        std::vector<double> e2 = rin;
        run_group(nt, [&](CoopMT co) { dct2_lines_pow2(co, e2.data(), e2.data(), a2.data(), b2.data(), n, nl, tw.data(), tw4.data()); });
        snprintf(name, sizeof name, "dct2_lines_pow2 in place n=%d lines=%d threads=%d", n, nl, nt);
        expect_close(name, e2.data(), d1.data(), d1.size(), 1e-13);
    } else {
        CoopHost one;
        std::vector<cplx> o1(in.size()), o2(in.size());
        dft_lines_direct(one, in.data(), o1.data(), n, nl, tw.data());
        run_group(nt, [&](CoopMT co) { dft_lines_direct(co, in.data(), o2.data(), n, nl, tw.data()); });
        snprintf(name, sizeof name, "dft_lines_direct n=%d lines=%d threads=%d", n, nl, nt);
        expect_close(name, flat(o2), flat(o1), 2 * in.size(), 1e-13);
        std::vector<double> d1(rin.size()), d2(rin.size());
        dct2_lines_direct(one, rin.data(), d1.data(), n, nl, c4.data());
        run_group(nt, [&](CoopMT co) { dct2_lines_direct(co, rin.data(), d2.data(), n, nl, c4.data()); });
        snprintf(name, sizeof name, "dct2_lines_direct n=%d lines=%d threads=%d", n, nl, nt);
        expect_close(name, d2.data(), d1.data(), d1.size(), 1e-13);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the whole 1D grid stage (k_kde1d's body)
// ---------------------------------------------------------------------------------------------------------------
struct Kde1dCase {
    int F;
    int bot, top, periodic, bco, mbc;
    double smooth;  // smooth_scale_1D (<= 0: automatic bandwidth)
    int likes;
};

struct Kde1dBuffers {
    std::vector<double> b, a2, logI, P, aux, aux2, win, raw, likes_out;
    std::vector<cplx> ca, cb, tw, tw4;
    std::vector<double> c4;
    Kde1dWork W{};
    Kde1dBuffers(int F, const std::vector<double>& bins, const double* likebins)
        : b(bins), a2(F), logI(F), P(F), aux(F), aux2(F), win(F + 1), raw(F), likes_out(F), ca(F), cb(F) {
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
        if (likebins) {
            W.likebins = likebins;
            W.raw = raw.data();
            W.likes_out = likes_out.data();
        }
    }
};

static void check_kde1d(const Kde1dCase& c, int nt) {
    const int F = c.F;
    // a histogram of a two-component mixture on [0, 1], piled up at the lower edge when that side is bounded
    std::vector<double> bins(F, 0.0), likebins(F, 0.0);
    const int N = 20000;
    for (int i = 0; i < N; i++) {
        double x = (urand() < 0.6) ? 0.35 + 0.08 * nrand() : 0.62 + 0.05 * nrand();
        if (c.bot) x = fabs(x - 0.3) + 0.05;
        if (c.periodic) x = x - floor(x);
        if (x < 0.02 || x > 0.98) continue;
        const int ix = (int)floor(x * (F - 1) + 0.5);
        const double w = 0.2 + urand();
        bins[ix] += w;
        likebins[ix] += w * exp(-2 * urand());
    }
    gdk_spec1d sp{};
    sp.param = 0;
    sp.fine_bins = F;
    sp.binmin = 0.0;
    sp.binmax = 1.0;
    sp.range_min = c.bot ? 0.05 : 0.1;
    sp.range_max = 0.9;
    sp.param_min = 0.06;
    sp.param_max = 0.88;
    sp.sigma_range = 0.12;
    sp.err = 0.14;
    sp.neff = 9000.0;
    sp.smooth_scale_1D = c.smooth;
    sp.width = (sp.range_max - sp.range_min) / 99;
    sp.boundary_correction_order = c.bco;
    sp.mult_bias_correction_order = c.mbc;
    sp.has_limits_bot = c.bot;
    sp.has_limits_top = c.top;
    sp.periodic = c.periodic;
    IsjConsts K;
    gdk_fill_isj_consts(&K);
    std::vector<double> P1(F), P2(F);
    gdk_result1d r1{}, r2{};
    Kde1dBuffers s1(F, bins, c.likes ? likebins.data() : nullptr), s2(F, bins, c.likes ? likebins.data() : nullptr);
    CoopHost one;
    kde1d_core(one, sp, K, s1.W, P1.data(), &r1);
    run_group(nt, [&](CoopMT co) {
        Kde1dWork W = s2.W;  // the pointers: per-thread copy, as the kernel's registers
        kde1d_core(co, sp, K, W, P2.data(), &r2);
    });
    char name[160];
    snprintf(name, sizeof name, "kde1d_core F=%d bot=%d top=%d per=%d bco=%d mbc=%d smooth=%g likes=%d thr=%d", F, c.bot, c.top,
             c.periodic, c.bco, c.mbc, c.smooth, c.likes, nt);
    expect_close(name, P2.data(), P1.data(), F, 1e-9);
    if (c.likes) expect_close("   ... likes", s2.likes_out.data(), s1.likes_out.data(), F, 1e-9);
    if (r1.winw != r2.winw || r1.status != r2.status) {
        printf("   result record differs: winw %d/%d status %u/%u\n", r2.winw, r1.winw, r2.status, r1.status);
        g_fail = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the 2D bandwidth selection (k_bw2d's body) and the contour levels (k_contours2d's body)
// ---------------------------------------------------------------------------------------------------------------
static void make_grid2d(int G, double rho, std::vector<double>& hist) {
    hist.assign((size_t)G * G, 0.0);
    const int N = 60000;
    for (int i = 0; i < N; i++) {
        const double u = nrand(), v = nrand();
        const double x = 0.5 + 0.09 * u, y = 0.5 + 0.09 * (rho * u + sqrt(1 - rho * rho) * v);
        if (x < 0.02 || x > 0.98 || y < 0.02 || y > 0.98) continue;
        const int ix = (int)floor(x * (G - 1) + 0.5), iy = (int)floor(y * (G - 1) + 0.5);
        hist[(size_t)iy * G + ix] += 0.2 + urand();
    }
}

// squared 2D DCT-II and |FFT2|^2 of hist / sum(hist) with one thread (inputs of the bandwidth stage)
static void xform2d(const std::vector<double>& hist, int G, std::vector<double>& a2, std::vector<double>& aF) {
    CoopHost co;
    const size_t n2 = (size_t)G * G;
    double total = 0;
    for (double v : hist) total += v;
    std::vector<double> p(n2), t1(n2), t2(n2);
    for (size_t i = 0; i < n2; i++) p[i] = hist[i] / total;
    std::vector<cplx> a(n2), b(n2), tw(G), c1(n2), c2(n2);
    std::vector<double> c4(4 * (size_t)G);
    gdk_fill_roots(tw.data(), G, G);
    gdk_fill_cos(c4.data(), 4 * G);
    dct2_lines_direct(co, p.data(), t1.data(), G, G, c4.data());
    for (int y = 0; y < G; y++)
        for (int x = 0; x < G; x++) t2[(size_t)x * G + y] = t1[(size_t)y * G + x];
    dct2_lines_direct(co, t2.data(), t1.data(), G, G, c4.data());
    a2.resize(n2);
    aF.resize(n2);
    for (int y = 0; y < G; y++)
        for (int x = 0; x < G; x++) {
            const double v = t1[(size_t)x * G + y];
            a2[(size_t)y * G + x] = v * v;
        }
    for (size_t i = 0; i < n2; i++) a[i] = cplx{p[i], 0};
    dft_lines_direct(co, a.data(), b.data(), G, G, tw.data());
    for (int y = 0; y < G; y++)
        for (int x = 0; x < G; x++) c1[(size_t)x * G + y] = b[(size_t)y * G + x];
    dft_lines_direct(co, c1.data(), c2.data(), G, G, tw.data());
    for (int y = 0; y < G; y++)
        for (int x = 0; x < G; x++) {
            const cplx v = c2[(size_t)x * G + y];
            aF[(size_t)y * G + x] = v.x * v.x + v.y * v.y;
        }
}

static void check_bw2d(int G, double rho, int do_corr, int nt) {
    std::vector<double> hist, a2, aF;
    make_grid2d(G, rho, hist);
    xform2d(hist, G, a2, aF);
    Kde2dConsts K;
    gdk_fill_kde2d_consts(&K);
    auto run = [&](auto runner) {
        std::vector<double> wx((size_t)PSI_MAXE * G), wy((size_t)PSI_MAXE * G);
        std::vector<int> cut(2 * PSI_MAXE);
        std::vector<PsiEntry> ebuf(PSI_MAXE);
        Kde2dWork W{a2.data(), do_corr ? aF.data() : nullptr, G, wx.data(), wy.data(), cut.data(), ebuf.data()};
        return runner(W);
    };
    CoopHost one;
    const Bw2dOut o1 = run([&](Kde2dWork& W) { return kernel_optimizer_2d(one, K, W, 30000.0, rho, do_corr, 0, 0.0); });
    Bw2dOut o2{};
    run([&](Kde2dWork& W) {
        run_group(nt, [&](CoopMT co) {
            const Bw2dOut o = kernel_optimizer_2d(co, K, W, 30000.0, rho, do_corr, 0, 0.0);
            if (co.tid == 0) o2 = o;
        });
        return 0;
    });
    char name[128];
    snprintf(name, sizeof name, "kernel_optimizer_2d G=%d rho=%.2f corr=%d threads=%d", G, rho, do_corr, nt);
    const double v1[4] = {o1.hx, o1.hy, o1.c, o1.t_star}, v2[4] = {o2.hx, o2.hy, o2.c, o2.t_star};
    expect_close(name, v2, v1, 4, 1e-6);
    if (o1.status != o2.status || o1.failed != o2.failed) {
        printf("   status differs: %u/%u failed %d/%d\n", o2.status, o1.status, o2.failed, o1.failed);
        g_fail = 1;
    }
}

static void check_contours(int G, int nt) {
    std::vector<double> P;
    make_grid2d(G, 0.5, P);
    // smooth a little so that the levels are well separated, then max-normalise (what k_finalize2d hands over)
    std::vector<double> Q(P.size());
    double mx = 0;
    for (int y = 0; y < G; y++)
        for (int x = 0; x < G; x++) {
            double acc = 0;
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++) {
                    const int yy = y + dy, xx = x + dx;
                    if (yy >= 0 && yy < G && xx >= 0 && xx < G) acc += P[(size_t)yy * G + xx] * exp(-0.3 * (dx * dx + dy * dy));
                }
            Q[(size_t)y * G + x] = acc;
            mx = fmax(mx, acc);
        }
    for (auto& v : Q) v /= mx;
    const double conts[3] = {0.68, 0.95, 0.99};
    double l1[4] = {0, 0, 0, 0}, l2[4] = {0, 0, 0, 0};
    CoopHost one;
    const unsigned s1 = contour_levels_core(one, Q.data(), G, conts, 3, l1);
    unsigned s2 = 0;
    run_group(nt, [&](CoopMT co) {
        double lv[4] = {0, 0, 0, 0};
        const unsigned s = contour_levels_core(co, Q.data(), G, conts, 3, lv);
        if (co.tid == 0) {
            s2 = s;
            for (int k = 0; k < 4; k++) l2[k] = lv[k];
        }
    });
    char name[128];
    snprintf(name, sizeof name, "contour_levels_core G=%d threads=%d", G, nt);
    expect_close(name, l2, l1, 3, 1e-12);
    if (s1 != s2) {
        printf("   status differs: %u/%u\n", s2, s1);
        g_fail = 1;
    }
}

int main(int argc, char** argv) {
    const int nt = argc > 1 ? atoi(argv[1]) : 64;  // a multiple of 32 (bilinear deals rows to warps)
    if (nt < 32 || nt % 32) {
        fprintf(stderr, "threads must be a multiple of 32\n");
        return 2;
    }
    if (argc > 2 && !strcmp(argv[2], "--break-barriers")) {
        // the checker checking itself: the FFT stages without their barriers are a textbook race
        g_break_barriers = true;
        check_transforms(64, 3, nt);
        return 0;  // ThreadSanitizer turns this into its own exit code when it has reported something
    }
    check_transforms(64, 3, nt);    // radix-4 stages only
    check_transforms(128, 2, nt);   // + the final radix-2 stage
    check_transforms(8, 5, nt);     // fewer butterflies than threads
    check_transforms(48, 2, nt);    // direct transforms (non power of two)
    const Kde1dCase cases[] = {
        {256, 0, 0, 0, 1, 1, -1.0, 0},   // automatic bandwidth, unbounded
        {256, 1, 0, 0, 1, 1, -1.0, 0},   // lower hard limit, linear boundary kernel
        {256, 1, 1, 0, 2, 1, -1.0, 0},   // both limits, quadratic boundary kernel
        {256, 1, 0, 0, 0, 0, -1.0, 0},   // normalised only
        {256, 0, 0, 0, 2, 2, -1.0, 0},   // higher-order kernel for an unbounded parameter, two bias iterations
        {256, 0, 0, 1, 1, 1, -1.0, 0},   // periodic
        {96, 0, 0, 0, 1, 1, -1.0, 0},    // direct DCT (F not a power of two)
        {256, 0, 1, 0, 1, 1, 0.3, 0},    // fixed smoothing in units of the standard deviation
        {256, 0, 0, 0, 1, 1, 2.0, 0},    // fixed smoothing in units of the coarse bin width
        {256, 1, 0, 0, 1, 1, -1.0, 1},   // mean likelihoods, bounded
        {256, 0, 0, 1, 1, 1, -1.0, 1},   // mean likelihoods, periodic
    };
    for (const auto& c : cases) check_kde1d(c, nt);
    check_bw2d(32, 0.0, 0, nt);
    check_bw2d(32, 0.6, 1, nt);
    check_bw2d(24, 0.4, 1, nt);
    check_contours(32, nt);
    check_contours(24, nt);
    printf(g_fail ? "FAILED\n" : "all cases agree\n");
    return g_fail;
}
