// kernels_2d.cuh -- the 2D density path on the device.
//   k_hist2d_tiles   weighted 2D fine-grid histograms (_binSamples x2 + _make2Dhist, mcsamples.py:1821-1827,
//                    1724-1728) for tiles of up to 8 x 8 parameter pairs per CTA: every sample row is read
//                    once per tile, its 16 bin indices computed once, then up to 64 grid updates issued as
//                    native 64-bit integer global reductions (REDG.ADD.64) into L2-resident grids
//   k_shear_*        the sheared re-binning of getAutoBandwidth2D (mcsamples.py:1370-1375, kde.bin_samples)
//   k_xform_*        dct2d and |fft2|^2 of the normalised histogram (kde_bandwidth.py:151-157) with the
//                    shared-memory Stockham FFT, rows then columns
//   k_bw2d           one CTA per pair: KernelOptimizer2D + the tail of getAutoBandwidth2D (kde2d_core.cuh)
//   k_build_kernel2d the correlated Gaussian window (mcsamples.py:1863-1867)
//   k_conv2d         direct 2D convolution, register-tiled (convolve2D 'same', convolve.py:205-212)
//   k_mask_*         the 'valid' convolutions of the prior mask with the moment kernels (mcsamples.py:
//                    1926-1951, 1967) evaluated separably: mask = my (x) mx is piecewise constant
//   k_boundary2d     linear boundary correction (mcsamples.py:1927-1959)
//   k_finalize2d     normalize('max') (densities.py:71-92)
#pragma once
#include <type_traits>
#include "kde2d_core.cuh"
#include "kernels_1d.cuh"

// ----------------------------------------------------------------------------------------------------
// histograms
// ----------------------------------------------------------------------------------------------------
#define HT 8  // tile edge (parameters per side)
static_assert(true, "");
#define HW 64 // edge of a shared-memory hot window (bins), see k_hist2d_hot / k_shear_hist (80 was measured: no gain)

struct Tile2d {
    int na, nb, G, pad;
    int pa[HT], pb[HT];
    double amin[HT], afw[HT], ainv[HT];
    double bmin[HT], bfw[HT], binv[HT];
    long long off[HT][HT];  // grid offset (elements) of pair (a, b) or -1
};

// grid (nseg, ntiles), 256 threads
__global__ void __launch_bounds__(256) k_hist2d_tiles(const double* __restrict__ dX, int64_t ld,
                                                      const unsigned long long* __restrict__ dWq, const Seg* __restrict__ segs,
                                                      const Tile2d* __restrict__ tiles, unsigned long long* __restrict__ grids) {
    __shared__ Tile2d T;
    {
        const int* src = reinterpret_cast<const int*>(tiles + blockIdx.y);
        int* dst = reinterpret_cast<int*>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(Tile2d) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const Seg sg = segs[blockIdx.x];
    const int G = T.G, na = T.na, nb = T.nb;
    for (int64_t r = sg.r0 + threadIdx.x; r < sg.r1; r += blockDim.x) {
        const unsigned long long w = dWq[r];
        int ia[HT], ib[HT];
#pragma unroll
        for (int k = 0; k < HT; k++) {
            ia[k] = -1;
            if (k < na) {
                const int b = bin_index_round(ldg_stream(dX + (int64_t)T.pa[k] * ld + r), T.amin[k], T.afw[k], T.ainv[k]);
                ia[k] = (b >= 0 && b < G) ? b : -1;
            }
        }
#pragma unroll
        for (int k = 0; k < HT; k++) {
            ib[k] = -1;
            if (k < nb) {
                const int b = bin_index_round(ldg_stream(dX + (int64_t)T.pb[k] * ld + r), T.bmin[k], T.bfw[k], T.binv[k]);
                ib[k] = (b >= 0 && b < G) ? b : -1;
            }
        }
        if (w == 0) continue;
#pragma unroll
        for (int a = 0; a < HT; a++) {
            if (a >= na || ia[a] < 0) continue;
#pragma unroll
            for (int b = 0; b < HT; b++) {
                if (b >= nb || ib[b] < 0) continue;
                const long long off = T.off[a][b];
                if (off >= 0) atomicAdd(grids + off + (long long)ib[b] * G + ia[a], w);
            }
        }
    }
}

// in-place fixed point -> float64 (same storage)
__global__ void k_u64_to_f64_inplace(unsigned long long* __restrict__ g, int64_t n, double inv_scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)g[i] * inv_scale;
        reinterpret_cast<double*>(g)[i] = v;
    }
}

struct ShearJob {
    int pi, pj, Gb, pad;
    double r0, r1;
    double p1_min, dx1, inv1;  // kde.bin_samples geometry of p1 (host)
    long long off;             // rot grid offset (elements)
};
struct ShearGeom {
    double rmin, dx, inv, R;
};

__device__ __forceinline__ double shear_p2(double xi, double xj, double r0, double r1) {
    return __dadd_rn(__dmul_rn(r0, xi), __dmul_rn(r1, xj));  // r[0]*xi + r[1]*xj, no FMA contraction
}

// Jobs that share their p1 column are processed together: x_i is read once per row for up to SG jobs.
#define SG 6  // max jobs per group (k_shear_hist_w<NP, W>: NP <= SG windows of W x W bins in shared memory)
struct ShearGroup {
    int pi, nj, Gb, pad;
    int pj[SG], job[SG];
    double r0[SG], r1[SG];
    double p1_min, dx1, inv1;
    long long off[SG];
    double mean1, mean2[SG];  // weighted means of p1 and of each p2 = r0*mean_i + r1*mean_j (window centres)
};

// grid (nseg, ngroups): partial min/max of p2 per job -> part[(job*nseg + seg)*2]
__global__ void __launch_bounds__(256) k_shear_minmax(const double* __restrict__ dX, int64_t ld, const Seg* __restrict__ segs,
                                                      int nseg, const ShearGroup* __restrict__ groups, double* __restrict__ part) {
    __shared__ ShearGroup g;
    __shared__ double sh[2][SG][8];
    {
        const int* src = reinterpret_cast<const int*>(groups + blockIdx.y);
        int* dst = reinterpret_cast<int*>(&g);
        for (int i = threadIdx.x; i < (int)(sizeof(ShearGroup) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const Seg sg = segs[blockIdx.x];
    const double* xi = dX + (int64_t)g.pi * ld;
    const int nj = g.nj;
    double mn[SG], mx[SG];
#pragma unroll
    for (int k = 0; k < SG; k++) {
        mn[k] = INFINITY;
        mx[k] = -INFINITY;
    }
    for (int64_t r = sg.r0 + threadIdx.x; r < sg.r1; r += blockDim.x) {
        const double a = ldg_stream(xi + r);
        double b[SG];
#pragma unroll
        for (int k = 0; k < SG; k++)
            if (k < nj) b[k] = ldg_stream(dX + (int64_t)g.pj[k] * ld + r);
#pragma unroll
        for (int k = 0; k < SG; k++)
            if (k < nj) {
                const double p2 = shear_p2(a, b[k], g.r0[k], g.r1[k]);
                mn[k] = fmin(mn[k], p2);
                mx[k] = fmax(mx[k], p2);
            }
    }
    const int wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < SG; k++) {
        const double a = warp_min(mn[k]), b = warp_max(mx[k]);
        if ((threadIdx.x & 31) == 0) {
            sh[0][k][wid] = a;
            sh[1][k][wid] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < nj) {
        const int k = threadIdx.x;
        double a = sh[0][k][0], b = sh[1][k][0];
        for (int i = 1; i < 8; i++) {
            a = fmin(a, sh[0][k][i]);
            b = fmax(b, sh[1][k][i]);
        }
        part[((int64_t)g.job[k] * nseg + blockIdx.x) * 2 + 0] = a;
        part[((int64_t)g.job[k] * nseg + blockIdx.x) * 2 + 1] = b;
    }
}

// kde.bin_samples range of p2 (kde_bandwidth.py:77-86): one thread per job
__global__ void k_shear_geom(const double* __restrict__ part, int nseg, int njobs, const ShearJob* __restrict__ jobs,
                             ShearGeom* __restrict__ geom) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    double mn = INFINITY, mx = -INFINITY;
    for (int s = 0; s < nseg; s++) {
        mn = fmin(mn, part[((int64_t)j * nseg + s) * 2]);
        mx = fmax(mx, part[((int64_t)j * nseg + s) * 2 + 1]);
    }
    const double delta = __dsub_rn(mx, mn);
    const double rmin = __dsub_rn(mn, __dmul_rn(delta, 0.1));
    const double rmax = __dadd_rn(mx, __dmul_rn(delta, 0.1));
    const double R = __dsub_rn(rmax, rmin);
    const double dx = __ddiv_rn(R, (double)(jobs[j].Gb - 1));
    geom[j] = ShearGeom{rmin, dx, 1.0 / dx, R};
}

// Sheared re-binning (kde.bin_samples of (p1, r0*x_i + r1*x_j), kde_bandwidth.py:76-87) with hot-window privatisation:
// grid (ngroups, nseg) -- GROUPS are the fast grid index, so the CTAs resident at one time sweep the same row
// segment and the columns they share (an anchor's x_i, the partners it shares with neighbouring anchors) come out
// of L2 instead of DRAM.  512 threads, one CTA per SM, dynamic smem = NP * (W*W + 4) * 8 bytes: every job of the group
// keeps the W x W bins around the centre of its sheared cloud in shared memory (two 32-bit limbs per bin, native
// ATOMS); the few samples outside the window go to L2 reductions.  NP = partners of the group (exactly; the host
// launches one instantiation per group size), W = window edge for that size (3 x 96^2, 4 x 80^2, 6 x 64^2 ... fill
// the 227 KB of an SM).  Per-partner constants live in registers, rows are read two at a time with 16-byte loads and
// prefetched one iteration ahead.
template <int NP, int W>
__global__ void __launch_bounds__(512, 1) k_shear_hist_w(const double* __restrict__ dX, int64_t ld,
                                                         const unsigned long long* __restrict__ dWq, const Seg* __restrict__ segs,
                                                         const ShearGroup* __restrict__ groups, const ShearGeom* __restrict__ geom,
                                                         unsigned long long* __restrict__ grids) {
    extern __shared__ unsigned ssm2[];  // per job: lo[WB], hi[WB], WB = W*W + 4 (bin W*W of job 0 is the trash bin)
    constexpr unsigned WB = W * W + 4;
    __shared__ ShearGroup g;
    __shared__ ShearGeom gm[SG];
    __shared__ int s_ax0, s_by0[SG];
    {
        const int* src = reinterpret_cast<const int*>(groups + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&g);
        for (int i = threadIdx.x; i < (int)(sizeof(ShearGroup) / 4); i += blockDim.x) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i < NP * 2 * (int)WB; i += blockDim.x) ssm2[i] = 0;
    __syncthreads();
    const int G = g.Gb;
    if (threadIdx.x < NP) {
        gm[threadIdx.x] = geom[g.job[threadIdx.x]];
        const int c = bin_index_trunc(g.mean2[threadIdx.x], gm[threadIdx.x].rmin, gm[threadIdx.x].dx, gm[threadIdx.x].inv) - W / 2;
        s_by0[threadIdx.x] = max(0, min(G - W, c));
    }
    if (threadIdx.x == 0) {
        const int c = bin_index_trunc(g.mean1, g.p1_min, g.dx1, g.inv1) - W / 2;
        s_ax0 = max(0, min(G - W, c));
    }
    __syncthreads();
    const unsigned wlim = G >= W ? (unsigned)W : 0u;  // grids smaller than the window: everything takes the L2 path
    unsigned smem0 = (unsigned)__cvta_generic_to_shared(ssm2);
    asm volatile("mov.u32 %0, %0;" : "+r"(smem0));  // opaque: stays in a register (see k_hist2d_hot)
    const Seg sg = segs[blockIdx.y];
    const double p1_min = g.p1_min, dx1 = g.dx1, inv1s = g.inv1 * 1048576.0;
    const int ax0 = s_ax0;
    double r0[NP], r1[NP], rmin[NP], invs[NP];
    int by0[NP];
    const double* xj[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
        r0[k] = g.r0[k];
        r1[k] = g.r1[k];
        rmin[k] = gm[k].rmin;
        invs[k] = gm[k].inv * 1048576.0;
        by0[k] = s_by0[k];
        xj[k] = dX + (int64_t)g.pj[k] * ld;
    }
    const double* xi = dX + (int64_t)g.pi * ld;
    unsigned long long* gk[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) gk[k] = grids + g.off[k];
    // Two rows (2 * NP updates) at a time, written for instruction-level parallelism: all bin indices first (branch
    // free; the exact division is a rare fix-up of the whole group), then all low-limb ATOMS back to back, then the
    // carries into the high limbs, then the few out-of-window samples as L2 reductions.
    const unsigned trash = smem0 + W * W * 4u;  // spare bin behind window 0: target of the out-of-window samples
    auto update2 = [&](double a0, double a1, unsigned long long w0, unsigned long long w1, const double (&bx)[NP], const double (&by)[NP],
                       bool second) {
        const double d10 = __dsub_rn(a0, p1_min), d11 = __dsub_rn(a1, p1_min);
        const unsigned I10 = __double2uint_rd(__dmul_rn(d10, inv1s)), I11 = __double2uint_rd(__dmul_rn(d11, inv1s));
        auto edge = [](unsigned I) { return ((I & 0xfffffu) - 8u) >= (0xfffffu - 15u); };  // within 8 * 2^-20 of a bin edge / saturated
        bool bad = edge(I10) | edge(I11);
        double d2[2][NP];
        unsigned I2[2][NP];
#pragma unroll
        for (int k = 0; k < NP; k++) {
            d2[0][k] = __dsub_rn(shear_p2(a0, bx[k], r0[k], r1[k]), rmin[k]);
            d2[1][k] = __dsub_rn(shear_p2(a1, by[k], r0[k], r1[k]), rmin[k]);
            I2[0][k] = __double2uint_rd(__dmul_rn(d2[0][k], invs[k]));
            I2[1][k] = __double2uint_rd(__dmul_rn(d2[1][k], invs[k]));
            bad |= edge(I2[0][k]) | edge(I2[1][k]);
        }
        int b1[2] = {(int)(I10 >> 20), (int)(I11 >> 20)};
        int b2[2][NP];
#pragma unroll
        for (int k = 0; k < NP; k++) {
            b2[0][k] = (int)(I2[0][k] >> 20);
            b2[1][k] = (int)(I2[1][k] >> 20);
        }
        if (__builtin_expect(bad, 0)) {  // some index sits on a bin edge: the exact IEEE division decides (kde_bandwidth.py:85-87)
            if (edge(I10)) b1[0] = bin_index_trunc_exact(d10, dx1);
            if (edge(I11)) b1[1] = bin_index_trunc_exact(d11, dx1);
#pragma unroll
            for (int k = 0; k < NP; k++) {
                if (edge(I2[0][k])) b2[0][k] = bin_index_trunc_exact(d2[0][k], gm[k].dx);
                if (edge(I2[1][k])) b2[1][k] = bin_index_trunc_exact(d2[1][k], gm[k].dx);
            }
        }
        // a row whose p1 falls outside the grid (samples beyond a hard limit) contributes nothing: weight 0, bin 0
        if ((unsigned)b1[0] >= (unsigned)G) {
            b1[0] = 0;
            w0 = 0;
        }
        if ((unsigned)b1[1] >= (unsigned)G || !second) {
            b1[1] = 0;
            w1 = 0;
        }
        const unsigned dxr[2] = {(unsigned)(b1[0] - ax0), (unsigned)(b1[1] - ax0)};
        const unsigned wl[2] = {(unsigned)w0, (unsigned)w1}, wh[2] = {(unsigned)(w0 >> 32), (unsigned)(w1 >> 32)};
        // sm_100a cannot predicate shared / global atomics (ptxas turns "@p atom" into a divergent branch with a
        // convergence barrier per update), so every lane issues its shared-memory update unconditionally: samples
        // outside the window are pointed at a trash bin behind the windows, and only their L2 reduction sits in a
        // (short) divergent region: one IMAD.WIDE + REDG.
        unsigned addr[2][NP], old[2][NP];
        int gi[2][NP];
        bool hit[2][NP];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < NP; k++) {
                const unsigned dy = (unsigned)(b2[r][k] - by0[k]);
                hit[r][k] = max(dxr[r], dy) < wlim;
                gi[r][k] = min(max(b2[r][k], 0), G - 1) * G + b1[r];  // p2 ranges cover the samples; the clamp only guards NaNs
                const unsigned in_win = smem0 + (unsigned)k * (2u * WB * 4u) + dy * (4u * W) + (dxr[r] << 2);
                addr[r][k] = hit[r][k] ? in_win : trash;
            }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < NP; k++)
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old[r][k]) : "r"(addr[r][k]), "r"(wl[r]));
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < NP; k++) {
                unsigned c;
                asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(c) : "r"(old[r][k]), "r"(wl[r]), "r"(wh[r]));
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr[r][k] + WB * 4u), "r"(c));
            }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int k = 0; k < NP; k++)
                if (!hit[r][k]) atomicAdd(gk[k] + gi[r][k], r ? w1 : w0);
    };
    // segments start on even rows (16-byte aligned pairs); an odd tail row is handled by thread 0
    const int64_t npair = (sg.r1 - sg.r0) >> 1;
    int64_t i = threadIdx.x;
    double2 a2, b2v[NP];
    ulonglong2 w2;
    if (i < npair) {
        a2 = ldg_stream2(xi + sg.r0 + 2 * i);
        w2 = ldg_stream2_u64(dWq + sg.r0 + 2 * i);
#pragma unroll
        for (int k = 0; k < NP; k++) b2v[k] = ldg_stream2(xj[k] + sg.r0 + 2 * i);
    }
    while (i < npair) {
        const int64_t in = i + blockDim.x;
        double2 an, bn[NP];
        ulonglong2 wn;
        if (in < npair) {  // next pair of rows in flight while this one is binned
            an = ldg_stream2(xi + sg.r0 + 2 * in);
            wn = ldg_stream2_u64(dWq + sg.r0 + 2 * in);
#pragma unroll
            for (int k = 0; k < NP; k++) bn[k] = ldg_stream2(xj[k] + sg.r0 + 2 * in);
        }
        {
            double bx[NP], by[NP];
#pragma unroll
            for (int k = 0; k < NP; k++) {
                bx[k] = b2v[k].x;
                by[k] = b2v[k].y;
            }
            update2(a2.x, a2.y, w2.x, w2.y, bx, by, true);
        }
        a2 = an;
        w2 = wn;
#pragma unroll
        for (int k = 0; k < NP; k++) b2v[k] = bn[k];
        i = in;
    }
    if (((sg.r1 - sg.r0) & 1) && threadIdx.x == 0) {
        const int64_t r = sg.r1 - 1;
        double bx[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) bx[k] = xj[k][r];
        update2(xi[r], xi[r], dWq[r], 0ull, bx, bx, false);
    }
    __syncthreads();
    if (!wlim) return;
#pragma unroll 1
    for (int k = 0; k < NP; k++) {
        const unsigned* base = ssm2 + k * 2 * WB;
        unsigned long long* gk = grids + g.off[k] + (long long)by0[k] * G + ax0;
        for (int t = threadIdx.x; t < W * W; t += blockDim.x) {
            const unsigned long long v = ((unsigned long long)base[WB + t] << 32) | base[t];
            if (v) atomicAdd(gk + (long long)(t / W) * G + (t % W), v);
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// transforms
// ----------------------------------------------------------------------------------------------------
struct XformJob {
    const double* src;  // G x G histogram (float64)
    double* a2;         // G x G
    double* aFFT;       // G x G or NULL
    double* tmpD;       // G x G real scratch
    cplx* tmpC;         // G x G complex scratch (NULL when aFFT is NULL)
    const cplx* tw;     // tables for G (pow2) ...
    const cplx* tw4;
    const double* cos4;  // ... or direct tables
    const cplx* twn;
    int G, pad;
    double total;  // filled by k_grid_totals
};

// one CTA per job
__global__ void __launch_bounds__(256) k_grid_totals(XformJob* __restrict__ jobs) {
    __shared__ double red[32];
    XformJob& jb = jobs[blockIdx.x];
    const int64_t n = (int64_t)jb.G * jb.G;
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += jb.src[i];
    CoopBlock co{(int)threadIdx.x, (int)blockDim.x, red};
    const double t = co.sum(s);
    if (threadIdx.x == 0) jb.total = t;
}

// rows: grid (ceil(G/rb), njobs), 256 threads, dynamic smem = rb*G*(8 + 16 + 16) bytes
__global__ void __launch_bounds__(256) k_xform_rows(const XformJob* __restrict__ jobs, int rb) {
    extern __shared__ __align__(16) unsigned char xsm[];
    __shared__ double red[32];
    const XformJob jb = jobs[blockIdx.y];
    const int G = jb.G;
    const int y0 = blockIdx.x * rb;
    if (y0 >= G) return;
    const int nl = min(rb, G - y0);
    double* in = reinterpret_cast<double*>(xsm);
    cplx* a = reinterpret_cast<cplx*>(xsm + (size_t)rb * G * 8);
    cplx* b = a + (size_t)rb * G;
    CoopBlock co{(int)threadIdx.x, (int)blockDim.x, red};
    for (int it = threadIdx.x; it < nl * G; it += blockDim.x) in[it] = jb.src[(size_t)y0 * G + it] / jb.total;
    __syncthreads();
    double* outD = jb.tmpD + (size_t)y0 * G;
    if (jb.tw) {
        dct2_lines_pow2(co, in, outD, a, b, G, nl, jb.tw, jb.tw4);
    } else {
        dct2_lines_direct(co, in, outD, G, nl, jb.cos4);
    }
    if (jb.aFFT) {
        cplx* outC = jb.tmpC + (size_t)y0 * G;
        if (jb.tw) {
            for (int it = threadIdx.x; it < nl * G; it += blockDim.x) a[it] = cplx{in[it], 0.0};
            __syncthreads();
            const cplx* r = fft_lines(co, a, b, G, nl, jb.tw);
            for (int it = threadIdx.x; it < nl * G; it += blockDim.x) outC[it] = r[it];
        } else {
            for (int it = threadIdx.x; it < nl * G; it += blockDim.x) a[it] = cplx{in[it], 0.0};
            __syncthreads();
            dft_lines_direct(co, a, outC, G, nl, jb.twn);
        }
    }
}

// columns: grid (ceil(G/cb), njobs), 256 threads, dynamic smem = cb*G*(8 + 8 + 16 + 16) bytes
__global__ void __launch_bounds__(256) k_xform_cols(const XformJob* __restrict__ jobs, int cb) {
    extern __shared__ __align__(16) unsigned char xsm[];
    __shared__ double red[32];
    const XformJob jb = jobs[blockIdx.y];
    const int G = jb.G;
    const int x0 = blockIdx.x * cb;
    if (x0 >= G) return;
    const int nl = min(cb, G - x0);
    double* in = reinterpret_cast<double*>(xsm);
    double* out = in + (size_t)cb * G;
    cplx* a = reinterpret_cast<cplx*>(xsm + (size_t)cb * G * 16);
    cplx* b = a + (size_t)cb * G;
    CoopBlock co{(int)threadIdx.x, (int)blockDim.x, red};
    // line l = column x0 + l, element y
    for (int it = threadIdx.x; it < nl * G; it += blockDim.x) {
        const int y = it / nl, l = it - y * nl;
        in[(size_t)l * G + y] = jb.tmpD[(size_t)y * G + x0 + l];
    }
    __syncthreads();
    if (jb.tw)
        dct2_lines_pow2(co, in, out, a, b, G, nl, jb.tw, jb.tw4);
    else
        dct2_lines_direct(co, in, out, G, nl, jb.cos4);
    for (int it = threadIdx.x; it < nl * G; it += blockDim.x) {
        const int y = it / nl, l = it - y * nl;
        const double v = out[(size_t)l * G + y];
        jb.a2[(size_t)y * G + x0 + l] = v * v;
    }
    if (jb.aFFT) {
        __syncthreads();
        for (int it = threadIdx.x; it < nl * G; it += blockDim.x) {
            const int y = it / nl, l = it - y * nl;
            a[(size_t)l * G + y] = jb.tmpC[(size_t)y * G + x0 + l];
        }
        __syncthreads();
        const cplx* r;
        if (jb.tw) {
            r = fft_lines(co, a, b, G, nl, jb.tw);
        } else {
            dft_lines_direct(co, a, b, G, nl, jb.twn);
            r = b;
        }
        for (int it = threadIdx.x; it < nl * G; it += blockDim.x) {
            const int y = it / nl, l = it - y * nl;
            const cplx v = r[(size_t)l * G + y];
            jb.aFFT[(size_t)y * G + x0 + l] = v.x * v.x + v.y * v.y;
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// bandwidth
// ----------------------------------------------------------------------------------------------------
struct Bw2dJob {
    const double* a2;
    const double* aFFT;
    int G;        // size of the optimiser grid (pair grid or base grid for the shear branch)
    int shear;    // index into ShearGeom or -1
};

// grid (npairs), 256 threads, dynamic smem = 2 * PSI_MAXE * Gmax * 8 + the row ring (COOP_RING_STAGES * ring_rows * Gmax * 8)
__global__ void __launch_bounds__(256, 2) k_bw2d(const gdk_spec2d* __restrict__ specs, const Bw2dJob* __restrict__ jobs,
                                                 const ShearGeom* __restrict__ geom, Kde2dConsts K, gdk_result2d* __restrict__ res,
                                                 int Gmax, int ring_rows) {
    extern __shared__ __align__(16) unsigned char bsm[];
    __shared__ double red[32], redv[8 * 32];
    __shared__ int cuts[2 * PSI_MAXE];
    __shared__ PsiEntry ebuf[PSI_MAXE];
    const gdk_spec2d sp = specs[blockIdx.x];
    const Bw2dJob jb = jobs[blockIdx.x];
    CoopBlock co{(int)threadIdx.x, (int)blockDim.x, red, redv,
                 ring_rows > 0 ? reinterpret_cast<double*>(bsm) + (size_t)2 * PSI_MAXE * Gmax : nullptr, ring_rows};
    Bw2dOut o{0, 0, 0, NAN, 0u, 0, 0};
    double r2 = 0;
    if (sp.bw_mode == GDK_BW2D_PLAIN || sp.bw_mode == GDK_BW2D_SHEAR) {
        const int G = jb.G;
        double* wx = reinterpret_cast<double*>(bsm);
        double* wy = wx + (size_t)PSI_MAXE * G;
        const bool has_limits = sp.x_has_bot || sp.x_has_top || sp.y_has_bot || sp.y_has_top;
        const int do_corr = has_limits ? 0 : 1;
        Kde2dWork W{jb.a2, do_corr ? jb.aFFT : nullptr, G, wx, wy, cuts, ebuf};
        if (sp.bw_mode == GDK_BW2D_PLAIN) {
            const double rangex = sp.xbinmax - sp.xbinmin, rangey = sp.ybinmax - sp.ybinmin;
            const double q = fmin(sp.y_sigma_range / rangey, sp.x_sigma_range / rangex) / pow(sp.neff, 1.0 / 6);
            o = kernel_optimizer_2d(co, K, W, sp.neff, sp.corr, do_corr, 1, q * q);
        } else {
            o = kernel_optimizer_2d(co, K, W, sp.neff, 0.0, do_corr, 0, 0.0);
            r2 = geom[jb.shear].R;
        }
    }
    if (threadIdx.x == 0) finish_bandwidth_2d(sp, &o, r2, res + blockIdx.x);
}

// ----------------------------------------------------------------------------------------------------
// convolution stage
// ----------------------------------------------------------------------------------------------------
struct ConvJob {
    const double* hist;  // G x G
    double* Wk;          // K x K window, [u + w][v + w]
    double* P;           // current density
    double* Pn;          // next density (bias iterations)
    double* xP;          // bounded, order 1
    double* yP;
    double* maps;        // bounded: 6 maps a00,a10,a01,a20,a02,a11 (G*G each)
    double* a00b;        // bias-correction normaliser map
    double* box;         // hist / P where P > 1e-8 max(P), else hist: input of the bias-correction convolution
    double* T;           // scratch K x G x 3
    unsigned long long* mx;  // [8] running maxima (bit patterns of non-negative doubles):
                             // 0 = hist*W, 1 = after boundary correction, 2 + it = after bias iteration it
    double rx, ry, c;
    int G, w;
    int bounded;  // boundary correction active (has_prior && order >= 0)
    int bco, mbc;
    int xb, xt, yb, yt;
    int nc, pad;
    double contours[4];
    int xper, yper;  // periodic axes
    // mean likelihoods (meanlikes=True, mcsamples.py:1829-1831, 1886-1901, 2004-2006); lhist == NULL: not requested
    const double* lhist;  // histogram of weights * exp(mean_loglike - loglikes)
    double* lP;           // lhist (*) Win
    double* lbox;         // lhist / lP where lP > 0
    double* lP2;          // conv(lbox) * lP where lP > 0 (bias-corrected likes); then reused for the ratio likes / bins2D
    int lmbc, pad3;       // the mult_bias_correction_order setting itself (the likes ignore periodicity, :1890)
    // user prior mask (mask_function, mcsamples.py:1909-1919): (G + 2w)^2 doubles or NULL
    const double* umask;
};

// window: one CTA per job
__global__ void __launch_bounds__(256) k_build_kernel2d(const ConvJob* __restrict__ jobs) {
    __shared__ double red[32];
    const ConvJob jb = jobs[blockIdx.x];
    const int w = jb.w, K = 2 * w + 1;
    // Cinv = inv([[ry^2, rx ry c], [rx ry c, rx^2]])  (mcsamples.py:1864)
    const double m00 = jb.ry * jb.ry, m01 = jb.rx * jb.ry * jb.c, m11 = jb.rx * jb.rx;
    const double det = m00 * m11 - m01 * m01;
    const double C00 = m11 / det, C11 = m00 / det, C10 = -m01 / det;
    CoopBlock co{(int)threadIdx.x, (int)blockDim.x, red};
    double part = 0;
    for (int it = threadIdx.x; it < K * K; it += blockDim.x) {
        const int i1 = it / K - w, i2 = it % K - w;
        const double v = exp(-((double)(i1 * i1) * C00 + (double)(i2 * i2) * C11 + 2 * C10 * (double)i1 * (double)i2) / 2);
        jb.Wk[it] = v;
        part += v;
    }
    const double s = co.sum(part);
    for (int it = threadIdx.x; it < K * K; it += blockDim.x) jb.Wk[it] = jb.Wk[it] / s;
}

__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* p, double v) {
    if (v > 0) atomicMax(p, (unsigned long long)__double_as_longlong(v));
}

#define CV_TY 16
#define CV_TX 128
#define CV_MX 8
#define CV_KC 8
// geometry of the shared-memory tiles for a launch whose largest half width is wmax
__host__ __device__ __forceinline__ int cv_kp(int wmax) { return ((2 * wmax + 1 + 7) & ~7) + 8; }  // padded kernel row
__host__ __device__ __forceinline__ int cv_q(int wmax) {
    int q = (CV_TX + ((2 * wmax + 1 + 7) & ~7)) / 8 + 2;
    while ((q & 3) != 2) q++;  // q = 2 (mod 4): the interleaved fill below is then bank-conflict free
    return q;
}
// Direct 'same' convolution out[y][x] = sum_{u,v} W(u,v) in[y-u][x-v], zero padded.
// MODE 0: in = hist                      -> P (and xP, yP when bounded with order 1); running max -> mx[0]
// MODE 1: in = box = hist / P where P > thr (thr = mx[1+iter]*1e-8) else hist
//                                         -> next P = P * conv / a00b; running max -> mx[2+iter]
// MODE 4: MODE 0 compiled without the moment maps (groups with no boundary correction of order 1): 24 accumulator
//         registers fewer, higher occupancy
// MODE 2: in = lhist (mean-likelihood histogram) -> lP           (jobs without lhist return)
// MODE 3: in = lbox                              -> lP2 = conv * lP where lP > 0, else conv   (jobs with lmbc only)
// grid (tiles, njobs), 256 threads = 16 (x) x 16 (y); every thread produces 8 consecutive outputs of one row.
// Shared-memory input tile: row r, column c stored at r*8q + (c&7)*q + (c>>3) ("8-way column interleave"): the
// sliding-window loads of the 16 threads of a row then hit 16 consecutive words (conflict free), and so do the
// coalesced fills.  Kernel rows are stored reversed and zero padded to a multiple of 8 taps; the window lives
// in 16 registers that rotate through a fully unrolled block of 8 taps (one 8-byte load per tap per thread).
template <int MODE>
__global__ void __launch_bounds__(256) k_conv2d(const ConvJob* __restrict__ jobs, int iter, int wmax, int kcmax) {
    extern __shared__ __align__(16) double csm[];
    const ConvJob jb = jobs[blockIdx.y];
    if (MODE == 1 && iter >= jb.mbc) return;
    if ((MODE == 2 || MODE == 3) && (!jb.lhist || (MODE == 3 && !jb.lmbc))) return;
    if (jb.xper | jb.yper) return;  // periodic pairs: k_conv2d_circ
    const int G = jb.G, w = jb.w, K = 2 * w + 1;
    const int Kp = (K + 7) & ~7;
    const int tilesx = (G + CV_TX - 1) / CV_TX;
    const int tyi = blockIdx.x / tilesx, txi = blockIdx.x % tilesx;
    const int oy0 = tyi * CV_TY, ox0 = txi * CV_TX;
    if (oy0 >= G) return;
    const int q = cv_q(wmax), pitch = 8 * q, kp = cv_kp(wmax);
    double* in_s = csm;
    double* wk_s = csm + (size_t)(CV_TY + kcmax - 1) * pitch;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double* P = (MODE == 1) ? ((iter & 1) ? jb.Pn : jb.P) : nullptr;
    double* Pout = (MODE == 1) ? ((iter & 1) ? jb.P : jb.Pn) : jb.P;
    double thr = 0;
    if (MODE == 1) thr = __longlong_as_double((long long)jb.mx[1 + iter]) * 1e-8;
    const bool moments = (MODE == 0) && jb.bounded && jb.bco == 1;  // MODE 4: MODE 0 for groups without moment maps
    double acc[CV_MX], accx[CV_MX], accy[CV_MX];
#pragma unroll
    for (int m = 0; m < CV_MX; m++) acc[m] = accx[m] = accy[m] = 0;
    const int ncol = CV_TX + Kp;  // columns b = ox0 - w + c, c < ncol (zero beyond the grid / beyond 2w)
    for (int k0 = 0; k0 < K; k0 += kcmax) {
        const int kc = min(kcmax, K - k0);
        __syncthreads();
        // kernel chunk, columns reversed and zero padded: wk_s[kk][kr] = W[k0+kk][2w - kr] for kr < K, else 0
        for (int it = threadIdx.x; it < kc * kp; it += blockDim.x) {
            const int kk = it / kp, kr = it - kk * kp;
            wk_s[it] = (kr < K) ? jb.Wk[(size_t)(k0 + kk) * K + (2 * w - kr)] : 0.0;
        }
        // input rows a = abase + r, r < CV_TY + kc - 1; one warp per row, lanes stride over the columns
        const int abase = oy0 - (k0 + kc - 1) + w;
        const int nrow = CV_TY + kc - 1;
        const double* srcp = (MODE == 1) ? jb.box : (MODE == 2 ? jb.lhist : (MODE == 3 ? jb.lbox : jb.hist));  // MODE 0, 4: hist
        for (int r = threadIdx.x >> 5; r < nrow; r += 8) {
            const int a = abase + r;
            const bool rowok = a >= 0 && a < G;
            const double* srow = srcp + (size_t)(rowok ? a : 0) * G;
            double* drow = in_s + r * pitch;
            for (int c = threadIdx.x & 31; c < pitch; c += 32) {
                const int b = ox0 - w + c;
                double v = 0;
                if (rowok && c < ncol + 8 && b >= 0 && b < G) v = srow[b];
                drow[(c & 7) * q + (c >> 3)] = v;
            }
        }
        __syncthreads();
        for (int kk = 0; kk < kc; kk++) {
            const double* irow = in_s + (ty + (kc - 1 - kk)) * pitch + tx;
            const double* wr = wk_s + kk * kp;
            const double u = (double)(k0 + kk - w);
            double v[CV_MX], nx[CV_MX];
#pragma unroll
            for (int m = 0; m < CV_MX; m++) v[m] = irow[m * q];  // columns tx*8 + m
            for (int tb = 0; tb < Kp; tb += 8) {
                const int qi = 1 + (tb >> 3);
#pragma unroll
                for (int j = 0; j < 8; j++) nx[j] = irow[j * q + qi];  // columns tx*8 + tb + 8 + j
                // eight taps as four 16-byte (broadcast) loads
                double wv8[8];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const double2 t2 = *reinterpret_cast<const double2*>(wr + tb + j);
                    wv8[j] = t2.x;
                    wv8[j + 1] = t2.y;
                }
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const double wv = wv8[j];
                    // window of tap tb+j: columns tx*8 + tb + j + m, m = 0..7  ->  v[j+m] for j+m < 8, else nx[j+m-8]
                    if (!moments) {
#pragma unroll
                        for (int m = 0; m < CV_MX; m++) acc[m] = fma(wv, (j + m < 8) ? v[j + m] : nx[j + m - 8], acc[m]);
                    } else {
                        const double wvx = wv * (double)(w - (tb + j));  // Win * indexes (column offset)
                        const double wvy = wv * u;                        // Win * y       (row offset)
#pragma unroll
                        for (int m = 0; m < CV_MX; m++) {
                            const double c = (j + m < 8) ? v[j + m] : nx[j + m - 8];
                            acc[m] = fma(wv, c, acc[m]);
                            accx[m] = fma(wvx, c, accx[m]);
                            accy[m] = fma(wvy, c, accy[m]);
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < CV_MX; m++) v[m] = nx[m];
            }
        }
    }
    const int oy = oy0 + ty;
    double tmax = 0;
    if (oy < G) {
#pragma unroll
        for (int m = 0; m < CV_MX; m++) {
            const int ox = ox0 + tx * CV_MX + m;
            if (ox >= G) continue;
            const size_t o = (size_t)oy * G + ox;
            if (MODE == 0 || MODE == 4) {
                jb.P[o] = acc[m];
                if (moments) {
                    jb.xP[o] = accx[m];
                    jb.yP[o] = accy[m];
                }
                tmax = fmax(tmax, acc[m]);
            } else if (MODE == 1) {
                // mask_function: masked pixels are not divided by the normaliser (mcsamples.py:1973-1976)
                const bool masked = jb.umask && jb.umask[(size_t)(oy + w) * (G + 2 * w) + ox + w] < 1e-8;
                const double vv = masked ? P[o] * acc[m] : P[o] * acc[m] / jb.a00b[o];
                Pout[o] = vv;
                tmax = fmax(tmax, vv);
            } else if (MODE == 2) {
                jb.lP[o] = acc[m];
            } else {
                const double l1 = jb.lP[o];
                jb.lP2[o] = l1 > 0 ? acc[m] * l1 : acc[m];
            }
        }
    }
    if (MODE == 2 || MODE == 3) return;
    tmax = warp_max(tmax);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(jb.mx + (MODE == 1 ? 2 + iter : 0), tmax);
}

// mean likelihoods, elementwise steps (mcsamples.py:1890-1898).  grid (64, njobs).
// STEP 0: lbox = lhist / lP where lP > 0 (jobs with lmbc)
// STEP 1: ratio = likes / bins2D where bins2D > 1e-4 max(bins2D), else 0 (bins2D = the first convolution, before any
//         correction: runs before k_boundary2d); stored in lP2, running max -> mx[15]
template <int STEP>
__global__ void __launch_bounds__(256) k_likes2d(const ConvJob* __restrict__ jobs) {
    const ConvJob jb = jobs[blockIdx.y];
    if (!jb.lhist || (STEP == 0 && !jb.lmbc)) return;
    const size_t gg = (size_t)jb.G * jb.G;
    const double thr = 1e-4 * __longlong_as_double((long long)jb.mx[0]);
    double tmax = 0;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < gg; o += (size_t)gridDim.x * blockDim.x) {
        if (STEP == 0) {
            const double l1 = jb.lP[o], h = jb.lhist[o];
            jb.lbox[o] = l1 > 0 ? h / l1 : h;
        } else {
            const double l = jb.lmbc ? jb.lP2[o] : jb.lP[o], p = jb.P[o];
            const double v = p > thr ? l / p : 0.0;
            jb.lP2[o] = v;
            tmax = fmax(tmax, v);
        }
    }
    if (STEP == 1) {
        tmax = warp_max(tmax);
        if ((threadIdx.x & 31) == 0) atomic_max_nonneg(jb.mx + 15, tmax);
    }
}

// max-normalised mean likelihoods into the output (mcsamples.py:2004-2006).  grid (64, njobs)
__global__ void __launch_bounds__(256) k_likes_finalize2d(const ConvJob* __restrict__ jobs, double* __restrict__ out,
                                                          const long long* __restrict__ offs) {
    const ConvJob jb = jobs[blockIdx.y];
    if (!jb.lhist) return;
    const size_t gg = (size_t)jb.G * jb.G;
    const double mx = __longlong_as_double((long long)jb.mx[15]);
    double* o = out + offs[blockIdx.y];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < gg; i += (size_t)gridDim.x * blockDim.x) o[i] = jb.lP2[i] / mx;
}

// prior-mask factors along one axis (mcsamples.py:1688-1712): index s relative to the unpadded grid
__device__ __forceinline__ double mask1d(int s, int G, int bot, int top) {
    if (s < 0) return bot ? 0.0 : 1.0;
    if (s > G - 1) return top ? 0.0 : 1.0;
    double m = 1.0;
    if (s == 0 && bot) m *= 0.5;
    if (s == G - 1 && top) m *= 0.5;
    return m;
}

// T[s][ku][x] = sum_v v^s W(u,v) mx(x - v)   for s = 0,1,2 (edge mask) and, at s = 3, the all-edge mask with
// s = 0.  grid (K, njobs), 256 threads.
__global__ void __launch_bounds__(256) k_mask_T(const ConvJob* __restrict__ jobs) {
    const ConvJob jb = jobs[blockIdx.y];
    const int G = jb.G, w = jb.w, K = 2 * w + 1;
    const int ku = blockIdx.x;
    if (ku >= K || !jb.T || jb.umask) return;
    const double* wrow = jb.Wk + (size_t)ku * K;
    const int eb = (jb.bounded && !jb.xper) ? jb.xb : 0, et = (jb.bounded && !jb.xper) ? jb.xt : 0;
    for (int x = threadIdx.x; x < G; x += blockDim.x) {
        double t0 = 0, t1 = 0, t2 = 0, tb = 0;
        for (int kv = 0; kv < K; kv++) {
            const int v = kv - w;
            const int s = x - v;
            const double wv = wrow[kv];
            const double m = mask1d(s, G, eb, et);
            const double wx = wv * (double)v;
            t0 += wv * m;
            t1 += wx * m;
            t2 += (wx * (double)v) * m;
            const double mb = (!jb.xper && (s < 0 || s > G - 1)) ? 0.0 : m;
            tb += wv * mb;
        }
        const size_t base = ((size_t)ku) * G + x;
        const size_t plane = (size_t)K * G;
        jb.T[base] = t0;
        jb.T[plane + base] = t1;
        jb.T[2 * plane + base] = t2;
        jb.T[3 * plane + base] = tb;
    }
}

// maps: a_rs[y][x] = sum_u u^r my(y - u) T_s[u][x].  grid (ceil(G/256), njobs), one thread per column x.
// Interior rows (w <= y <= G-1-w) see my == 1 for every tap, so each map equals the column's FULL sum there; only
// the <= 2w boundary rows need corrections, and only for the taps that reach the edge: O(K + w^2) work per column
// instead of O(G K).
__global__ void __launch_bounds__(256) k_mask_maps(const ConvJob* __restrict__ jobs) {
    const ConvJob jb = jobs[blockIdx.y];
    const int G = jb.G, w = jb.w, K = 2 * w + 1;
    const size_t plane = (size_t)K * G, gg = (size_t)G * G;
    if (!jb.T || jb.umask) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= G) return;
    const int eb = (jb.bounded && !jb.yper) ? jb.yb : 0, et = (jb.bounded && !jb.yper) ? jb.yt : 0;
    const double* T0 = jb.T + x;
    const double* T1 = jb.T + plane + x;
    const double* T2 = jb.T + 2 * plane + x;
    const double* T3 = jb.T + 3 * plane + x;
    double F00 = 0, F10 = 0, F01 = 0, F20 = 0, F02 = 0, F11 = 0, FB = 0;
    for (int ku = 0; ku < K; ku++) {
        const double du = (double)(ku - w);
        const size_t o = (size_t)ku * G;
        FB += T3[o];
        if (jb.bounded) {
            const double t0 = T0[o], t1 = T1[o];
            F00 += t0;
            F10 += t1;
            F20 += T2[o];
            F01 += du * t0;
            F02 += (du * du) * t0;
            F11 += du * t1;
        }
    }
    for (int y = 0; y < G; y++) {
        double a00 = F00, a10 = F10, a01 = F01, a20 = F20, a02 = F02, a11 = F11, ab = FB;
        // taps reaching the lower edge (s = y - u <= 0) and the upper edge (s >= G-1)
        for (int pass = 0; pass < 2; pass++) {
            int ulo, uhi;
            if (pass == 0) {
                ulo = y > -w ? y : -w;
                uhi = w;
            } else {
                ulo = -w;
                uhi = (y - (G - 1)) < w ? (y - (G - 1)) : w;
                if (uhi >= y && y <= w) uhi = y - 1;  // never double count (only possible when K > G)
            }
            for (int u = ulo; u <= uhi; u++) {
                const int sidx = y - u;
                const double m = mask1d(sidx, G, eb, et);
                const double mb = (!jb.yper && (sidx < 0 || sidx > G - 1)) ? 0.0 : m;
                const size_t o = (size_t)(u + w) * G;
                ab += (mb - 1.0) * T3[o];
                if (jb.bounded && m != 1.0) {
                    const double d = m - 1.0, du = (double)u;
                    const double t0 = T0[o], t1 = T1[o];
                    a00 += d * t0;
                    a10 += d * t1;
                    a20 += d * T2[o];
                    a01 += (d * du) * t0;
                    a02 += (d * du * du) * t0;
                    a11 += (d * du) * t1;
                }
            }
        }
        const size_t o = (size_t)y * G + x;
        if (jb.a00b) jb.a00b[o] = ab;
        if (jb.bounded) {
            jb.maps[o] = a00;
            jb.maps[gg + o] = a10;
            jb.maps[2 * gg + o] = a01;
            jb.maps[3 * gg + o] = a20;
            jb.maps[4 * gg + o] = a02;
            jb.maps[5 * gg + o] = a11;
        }
    }
}

// mask_function (mcsamples.py:1909-1919): the prior mask is an arbitrary (G + 2w)^2 array, so the six boundary-kernel
// moment maps and the bias normaliser are plain 'valid' convolutions of (user mask x edge masks) with the window
// moments; one thread per output pixel (rare path).  grid (ceil(G*G/256), njobs).
__global__ void __launch_bounds__(256) k_umask_maps(const ConvJob* __restrict__ jobs) {
    const ConvJob jb = jobs[blockIdx.y];
    if (!jb.umask) return;
    const int G = jb.G, w = jb.w, K = 2 * w + 1, Gp = G + 2 * w;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= G * G) return;
    const int y = o / G, x = o - y * G;
    const int xb = jb.bounded ? jb.xb : 0, xt = jb.bounded ? jb.xt : 0, yb = jb.bounded ? jb.yb : 0, yt = jb.bounded ? jb.yt : 0;
    double a00 = 0, a10 = 0, a01 = 0, a20 = 0, a02 = 0, a11 = 0, ab = 0;
    for (int ku = 0; ku < K; ku++) {
        const int u = ku - w, sy = y - u;
        const double my = mask1d(sy, G, yb, yt);
        const double myb = (sy < 0 || sy > G - 1) ? 0.0 : my;
        const double* urow = jb.umask + (size_t)(sy + w) * Gp + w;
        const double* wrow = jb.Wk + (size_t)ku * K;
        for (int kv = 0; kv < K; kv++) {
            const int v = kv - w, sx = x - v;
            const double um = urow[sx];
            const double mxe = mask1d(sx, G, xb, xt);
            const double m = um * mxe * my;
            const double mb = um * ((sx < 0 || sx > G - 1) ? 0.0 : mxe) * myb;
            const double wv = wrow[kv], wx = wv * (double)v, wy = wv * (double)u;
            a00 += wv * m;
            a10 += wx * m;
            a01 += wy * m;
            a20 += (wx * (double)v) * m;
            a02 += (wy * (double)u) * m;
            a11 += (wy * (double)v) * m;
            ab += wv * mb;
        }
    }
    const size_t gg = (size_t)G * G;
    if (jb.a00b) jb.a00b[o] = ab;
    if (jb.bounded) {
        jb.maps[o] = a00;
        jb.maps[gg + o] = a10;
        jb.maps[2 * gg + o] = a01;
        jb.maps[3 * gg + o] = a20;
        jb.maps[4 * gg + o] = a02;
        jb.maps[5 * gg + o] = a11;
    }
}

// bins2D[bool_mask] = 0 (mcsamples.py:1978-1979) and the maximum of what remains -> mx[14].  grid (64, njobs)
__global__ void __launch_bounds__(256) k_umask_zero(const ConvJob* __restrict__ jobs) {
    const ConvJob jb = jobs[blockIdx.y];
    if (!jb.umask) return;
    const int G = jb.G, w = jb.w;
    const size_t gg = (size_t)G * G;
    double* P = (jb.mbc & 1) ? jb.Pn : jb.P;
    double tmax = 0;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < gg; o += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(o / G), x = (int)(o - (size_t)y * G);
        if (jb.umask[(size_t)(y + w) * (G + 2 * w) + x + w] < 1e-8) P[o] = 0;
        tmax = fmax(tmax, P[o]);
    }
    tmax = warp_max(tmax);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(jb.mx + 14, tmax);
}

// boundary correction (mcsamples.py:1927-1959); running max of the corrected density -> mx[1]
__global__ void __launch_bounds__(256) k_boundary2d(const ConvJob* __restrict__ jobs) {
    const ConvJob jb = jobs[blockIdx.y];
    const int G = jb.G;
    const size_t gg = (size_t)G * G;
    const double mx0 = __longlong_as_double((long long)jb.mx[0]);
    double tmax = 0;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < gg; o += (size_t)gridDim.x * blockDim.x) {
        double p = jb.P[o];
        if (jb.bounded) {
            const double a00 = jb.maps[o];
            if (a00 * p > mx0 * 1e-8) {
                const double normed = p / a00;
                if (jb.bco == 0) {
                    p = normed;
                } else {
                    const double a10 = jb.maps[gg + o], a01 = jb.maps[2 * gg + o], a20 = jb.maps[3 * gg + o];
                    const double a02 = jb.maps[4 * gg + o], a11 = jb.maps[5 * gg + o];
                    const double xP = jb.xP[o], yP = jb.yP[o];
                    const double denom = a20 * a01 * a01 + a10 * a10 * a02 - a00 * a02 * a20 + a11 * a11 * a00 - 2 * a01 * a10 * a11;
                    const double A = a11 * a11 - a02 * a20;
                    const double Ax = a10 * a02 - a01 * a11;
                    const double Ay = a01 * a20 - a10 * a11;
                    const double corrected = (p * A + xP * Ax + yP * Ay) / denom;
                    p = normed * exp(fmin(corrected / normed, 4.0) - 1);
                }
                jb.P[o] = p;
            }
        }
        tmax = fmax(tmax, p);
    }
    tmax = warp_max(tmax);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(jb.mx + 1, tmax);
}

// normalize('max') into the output; status ZERO_MAX if the maximum is 0
__global__ void __launch_bounds__(256) k_finalize2d(const ConvJob* __restrict__ jobs, double* __restrict__ out,
                                                    const long long* __restrict__ offs, gdk_result2d* __restrict__ res) {
    const ConvJob jb = jobs[blockIdx.y];
    const int G = jb.G;
    const size_t gg = (size_t)G * G;
    // final density buffer and its max slot: after mbc bias iterations
    const int it = jb.mbc;
    const double* P = (it & 1) ? jb.Pn : jb.P;
    const int slot = jb.umask ? 14 : 1 + it;
    const double mx = __longlong_as_double((long long)jb.mx[slot]);
    double* o = out + offs[blockIdx.y];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < gg; i += (size_t)gridDim.x * blockDim.x)
        o[i] = (mx != 0) ? P[i] / mx : P[i];
    if (blockIdx.x == 0 && threadIdx.x == 0 && !(mx != 0)) res[blockIdx.y].status |= GDK_ST_ZERO_MAX;
}

// contour levels of the normalised output grids (densities.py:19-56).  grid (njobs), 512 threads.
// The reference sorts the grid and walks the cumulative sum; here the crossing value of every contour is found by a
// multi-pass radix selection over the IEEE bit patterns (non-negative doubles order like their bits): each pass
// histograms the halved-edge bin contents -- as 64-bit fixed point, 2^61 = the grid total, so the sums are exact
// integers and independent of the order of the shared-memory atomics -- of the elements inside the contour's current
// key interval by their leading CT_BITS bits, scans the histogram and narrows the interval to one digit: at most six
// sweeps of the (L2-resident) grid instead of one per key bit.  contour_finish then interpolates as the reference does.
#define CT_BITS 11
#define CT_NB (1 << CT_BITS)
#define CT_SMEM (4 * 2 * CT_NB * 4)
__global__ void __launch_bounds__(512) k_contours2d(const ConvJob* __restrict__ jobs, const double* __restrict__ out,
                                                     const long long* __restrict__ offs, gdk_result2d* __restrict__ res) {
    extern __shared__ unsigned ct_hist[];  // [4 contours][lo limbs CT_NB | hi limbs CT_NB]
    __shared__ double red[32];
    __shared__ unsigned long long s_klo[4], s_khi[4], s_below[4], s_target[4];
    __shared__ int s_shift[4], s_done[4];
    const ConvJob* jp = jobs + blockIdx.x;
    const int G = jp->G, n = G * G, nc = jp->nc;
    if (nc <= 0) return;
    CoopBlock co{(int)threadIdx.x, (int)blockDim.x, red};
    const double* P = out + offs[blockIdx.x];
    double part = 0, pmx = 0;
    for (int i = co.tid; i < n; i += co.nt) {
        const double v = P[i];
        part += v * edge_factor(i / G, i % G, G);
        pmx = fmax(pmx, v);
    }
    const double norm = co.sum(part);
    const double vmax = co.max(pmx);
    double target[4];
    for (int c = 0; c < 4; c++) target[c] = c < nc ? (1 - jp->contours[c]) * norm : 0;
    const double scale = norm > 0 ? 2305843009213693952.0 / norm : 0.0;  // 2^61 / total
    // first interval: the top 20 octaves below the maximum (contour levels of a density normalised to its maximum sit
    // there); what lies below is summed per thread, not histogrammed -- the tails of the grid would otherwise pile
    // onto a handful of exponent bins.  A contour whose level is lower still falls back to [0, klo0) afterwards.
    __shared__ unsigned long long s_below0;
    const unsigned long long klo0 = dbl_bits(vmax * 9.5367431640625e-07);  // vmax * 2^-20 (0 if vmax is tiny)
    if (threadIdx.x < 4) {
        const int c = threadIdx.x;
        s_klo[c] = klo0;
        s_khi[c] = dbl_bits(vmax);
        s_below[c] = 0;
        const double tq = target[c] * scale;
        s_target[c] = tq >= 1.0 ? (unsigned long long)__double2ull_ru(tq) : 1ull;
        s_shift[c] = q_shift_for(s_khi[c] - klo0, CT_BITS);
        s_done[c] = (c >= nc || s_khi[c] == 0) ? 1 : 0;
        if (c == 0) s_below0 = 0;
    }
    __syncthreads();
    for (int pass = 0; pass < 8; pass++) {
        const bool shared_hist = pass == 0;  // all contours still share [0, key(max)]: one histogram
        bool active = false;
        for (int c = 0; c < nc; c++) active = active || !s_done[c];
        if (!active) break;
        for (int i = threadIdx.x; i < (shared_hist ? 1 : nc) * 2 * CT_NB; i += blockDim.x) ct_hist[i] = 0;
        unsigned long long klo[4], khi[4];
        int shift[4], live[4];
        for (int c = 0; c < 4; c++) {
            klo[c] = s_klo[c];
            khi[c] = s_khi[c];
            shift[c] = s_shift[c];
            live[c] = c < nc && !s_done[c];
        }
        __syncthreads();
        unsigned long long qbelow = 0;
        for (int i = co.tid; i < n; i += co.nt) {
            const double v = P[i];
            const unsigned long long kb = dbl_bits(v);
            const unsigned long long q = __double2ull_rn(v * edge_factor(i / G, i % G, G) * scale);
            if (q == 0) continue;
            if (shared_hist) {
                if (kb < klo0) {
                    qbelow += q;
                } else {
                    const unsigned bin = (unsigned)((kb - klo0) >> shift[0]);
                    smem_add_u64(ct_hist + bin, ct_hist + CT_NB + bin, q);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (live[c] && kb >= klo[c] && kb <= khi[c]) {
                        const unsigned bin = (unsigned)((kb - klo[c]) >> shift[c]);
                        smem_add_u64(ct_hist + c * 2 * CT_NB + bin, ct_hist + c * 2 * CT_NB + CT_NB + bin, q);
                    }
            }
        }
        if (shared_hist) {
            qbelow = warp_sum_u64(qbelow);
            if ((threadIdx.x & 31) == 0 && qbelow) atomicAdd(&s_below0, qbelow);
        }
        __syncthreads();
        // warp c scans the histogram of contour c: lane l owns bins [64 l, 64 l + 64)
        const int wc = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (shared_hist && wc < nc && live[wc] && s_target[wc] <= s_below0) {
            // the level lies below the first interval: continue on [0, klo0) with the contour's own histogram
            if (lane == 0) {
                s_klo[wc] = 0;
                s_khi[wc] = klo0 - 1;
                s_shift[wc] = q_shift_for(klo0 - 1, CT_BITS);
            }
        } else if (wc < nc && live[wc]) {
            if (shared_hist && lane == 0) s_below[wc] = s_below0;
            __syncwarp();
            const unsigned* hl = ct_hist + (shared_hist ? 0 : wc * 2 * CT_NB);
            const unsigned* hh = hl + CT_NB;
            const int per = CT_NB / 32;
            unsigned long long mine = 0;
            int last_nonempty = -1;
            for (int k = 0; k < per; k++) {
                const unsigned long long hv = ((unsigned long long)hh[lane * per + k] << 32) | hl[lane * per + k];
                mine += hv;
                if (hv) last_nonempty = lane * per + k;
            }
            unsigned long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const unsigned long long before = incl - mine;  // bins of the lower lanes
            const unsigned long long need = s_target[wc] - s_below[wc];  // > 0 by construction
            // the lane whose range contains the crossing: first lane with inclusive sum >= need
            const unsigned hit = __ballot_sync(0xffffffffu, incl >= need);
            int lastbin = last_nonempty;
#pragma unroll
            for (int o = 16; o; o >>= 1) lastbin = max(lastbin, __shfl_xor_sync(0xffffffffu, lastbin, o));
            int bsel = -1;
            unsigned long long bbelow = 0;
            if (hit) {
                const int src = __ffs(hit) - 1;
                if (lane == src) {
                    unsigned long long run = before;
                    for (int k = 0; k < per; k++) {
                        const unsigned long long hv = ((unsigned long long)hh[lane * per + k] << 32) | hl[lane * per + k];
                        if (run + hv >= need) {
                            bsel = lane * per + k;
                            bbelow = run;
                            break;
                        }
                        run += hv;
                    }
                }
                bsel = __shfl_sync(0xffffffffu, bsel, src);
                bbelow = __shfl_sync(0xffffffffu, bbelow, src);
            } else {  // rounding left the total a hair under the target: the crossing is the largest element
                bsel = lastbin;
                const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
                const unsigned long long hv = bsel >= 0 ? (((unsigned long long)hh[bsel] << 32) | hl[bsel]) : 0;
                bbelow = total - hv;
            }
            if (lane == 0) {
                if (bsel < 0) {
                    s_done[wc] = 1;  // empty interval (all-zero grid): leave klo
                } else {
                    const unsigned long long nlo = s_klo[wc] + ((unsigned long long)bsel << s_shift[wc]);
                    unsigned long long nhi = nlo + ((1ull << s_shift[wc]) - 1);
                    if (nhi > s_khi[wc]) nhi = s_khi[wc];
                    s_below[wc] += bbelow;
                    s_klo[wc] = nlo;
                    s_khi[wc] = nhi;
                    if (s_shift[wc] == 0 || nlo == nhi) {
                        s_done[wc] = 1;
                        s_khi[wc] = nlo;
                    } else {
                        s_shift[wc] = q_shift_for(nhi - nlo, CT_BITS);
                    }
                }
            }
        }
        __syncthreads();
    }
    // the interval of a finished contour is one digit wide: its largest member key present in the grid is the crossing
    // key (shift 0: exactly the key).  A contour stopped by `nlo == nhi` or shift 0 has klo == the key.
    unsigned long long Kc[4];
    for (int c = 0; c < 4; c++) Kc[c] = s_klo[c];
    double lv[4] = {0, 0, 0, 0};
    const unsigned outside = contour_finish(co, P, G, Kc, target, nc, lv);
    if (threadIdx.x == 0) {
        for (int c = 0; c < 4; c++) res[blockIdx.x].levels[c] = lv[c];
        if (outside) res[blockIdx.x].status |= GDK_ST_CONTOUR_RANGE;
    }
}

// ----------------------------------------------------------------------------------------------------
// shared-memory privatised 2D histograms for the common 256 x 256 grids
// ----------------------------------------------------------------------------------------------------
// k_bin8: per-parameter bin indices as bytes (the 2D grid geometry of a parameter does not depend on its
// partner, mcsamples.py:1821-1822), so the pair sweeps below read 1 byte per coordinate instead of 8.
// grid (nseg, nparams), 256 threads.
struct Bin8Job {
    int param, pad;
    double binmin, fw, inv;
};
__global__ void __launch_bounds__(256) k_bin8(const double* __restrict__ dX, int64_t ld, const Seg* __restrict__ segs,
                                              const Bin8Job* __restrict__ jobs, unsigned char* __restrict__ out, int64_t old) {
    const Bin8Job jb = jobs[blockIdx.y];
    const Seg sg = segs[blockIdx.x];
    const double* x = dX + (int64_t)jb.param * ld;
    unsigned char* o = out + (int64_t)blockIdx.y * old;
    for (int64_t r = sg.r0 + threadIdx.x; r < sg.r1; r += blockDim.x) {
        const int b = bin_index_round(ldg_stream(x + r), jb.binmin, jb.fw, jb.inv);
        o[r] = (unsigned char)(b < 0 ? 0 : (b > 255 ? 255 : b));
    }
}

// k_hist2d_bands: one CLUSTER of 4 CTAs per parameter pair.  CTA `rank` owns the rows y with (y & 3) == rank of
// the 256 x 256 grid: 64 x 256 bins x two 32-bit limbs = 128 KB of shared memory, updated with native ATOMS.ADD
// (no L2 atomics: profiles/r1g shows the tiled REDG kernel at 81 % of L2 throughput).  The pre-binned sample
// stream (x byte, y byte, 64-bit fixed-point weight) is fetched ONCE per cluster: the leader CTA issues TMA bulk
// copies with .multicast::cluster into a 2-stage ring at the same shared-memory offsets of all four CTAs, each
// CTA's own `full` mbarrier receives the transaction bytes; when a CTA has consumed a stage it arrives (remote,
// release.cluster) on the leader's `empty` barrier.  Every CTA tests all samples of the stream (cheap byte test)
// and accumulates the quarter that falls in its rows; at the end it writes its rows of the grid with plain
// stores -- it is their only writer.
#define HB_CHUNK 4096
#define HB_STAGES 2
#define HB_THREADS 512
struct BandJob {
    const unsigned char* ia;   // x bins, N bytes
    const unsigned char* ib;   // y bins
    unsigned long long* grid;  // 256 x 256, [y][x], fixed point
};

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mcast(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar,
                                               unsigned short mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned long long* local_bar, unsigned target_rank) {
    unsigned raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"((unsigned)__cvta_generic_to_shared(local_bar)), "r"(target_rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(phase)
        : "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(HB_THREADS, 1)
    k_hist2d_bands(const BandJob* __restrict__ jobs, const unsigned long long* __restrict__ dWq, int64_t N) {
    extern __shared__ __align__(128) unsigned char bsm2[];
    __shared__ __align__(8) unsigned long long full[HB_STAGES], empty[HB_STAGES];
    const unsigned rank = cluster_ctarank();
    const BandJob jb = jobs[blockIdx.x >> 2];
    unsigned* hlo = reinterpret_cast<unsigned*>(bsm2);  // [64][256]
    unsigned* hhi = hlo + 64 * 256;
    unsigned char* stage0 = bsm2 + 2 * 64 * 256 * 4;
    const size_t stage_bytes = (size_t)HB_CHUNK * 10;  // x bytes | y bytes | weights
    for (int i = threadIdx.x; i < 2 * 64 * 256; i += blockDim.x) hlo[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < HB_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();  // every CTA's barriers exist before any multicast or remote arrive
    const int nchunks = (int)((N + HB_CHUNK - 1) / HB_CHUNK);
    auto chunk_cnt = [&](int c) { return (int)min((int64_t)HB_CHUNK, N - (int64_t)c * HB_CHUNK); };
    auto chunk_bytes = [&](int c) {
        const unsigned cnt = (unsigned)chunk_cnt(c);
        const unsigned b1 = (cnt + 15u) & ~15u;  // byte streams, padded to 16 (arrays are padded)
        return 2u * b1 + ((cnt + 1u) & ~1u) * 8u;
    };
    auto issue = [&](int c) {  // leader only
        const int s = c % HB_STAGES;
        const unsigned cnt = (unsigned)chunk_cnt(c);
        const unsigned b1 = (cnt + 15u) & ~15u;
        const int64_t off = (int64_t)c * HB_CHUNK;
        unsigned char* st = stage0 + (size_t)s * stage_bytes;
        bulk_g2s_mcast(st, jb.ia + off, b1, &full[s], (unsigned short)0xF);
        bulk_g2s_mcast(st + HB_CHUNK, jb.ib + off, b1, &full[s], (unsigned short)0xF);
        bulk_g2s_mcast(st + 2 * HB_CHUNK, dWq + off, ((cnt + 1u) & ~1u) * 8u, &full[s], (unsigned short)0xF);
    };
    if (threadIdx.x == 0) {
        for (int c = 0; c < HB_STAGES && c < nchunks; c++) mbar_expect_tx(&full[c % HB_STAGES], chunk_bytes(c));
        if (rank == 0)
            for (int c = 0; c < HB_STAGES && c < nchunks; c++) issue(c);
    }
    for (int c = 0; c < nchunks; c++) {
        const int s = c % HB_STAGES;
        const unsigned ph = (unsigned)((c / HB_STAGES) & 1);
        mbar_wait(&full[s], ph);
        const int cnt = chunk_cnt(c);
        const unsigned char* st = stage0 + (size_t)s * stage_bytes;
        const uint2* pa = reinterpret_cast<const uint2*>(st);
        const uint2* pb = reinterpret_cast<const uint2*>(st + HB_CHUNK);
        const unsigned long long* pw = reinterpret_cast<const unsigned long long*>(st + 2 * HB_CHUNK);
        // 512 threads x 8 consecutive rows = one chunk
        const int r0 = threadIdx.x * 8;
        if (r0 < cnt) {
            const uint2 a8 = pa[threadIdx.x], b8 = pb[threadIdx.x];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const unsigned a = ((k < 4 ? a8.x : a8.y) >> (8 * (k & 3))) & 0xffu;
                const unsigned b = ((k < 4 ? b8.x : b8.y) >> (8 * (k & 3))) & 0xffu;
                if ((b & 3u) == rank && r0 + k < cnt) {
                    const unsigned bin = ((b >> 2) << 8) | a;
                    smem_add_u64(hlo + bin, hhi + bin, pw[r0 + k]);
                }
            }
        }
        __syncthreads();  // everyone is done reading this stage
        if (threadIdx.x == 0) {
            mbar_arrive_remote(&empty[s], 0);  // tell the leader
            if (c + HB_STAGES < nchunks) mbar_expect_tx(&full[s], chunk_bytes(c + HB_STAGES));
            if (rank == 0 && c + HB_STAGES < nchunks) {
                mbar_wait_cluster(&empty[s], ph);
                issue(c + HB_STAGES);
            }
        }
    }
    __syncthreads();
    // write this CTA's rows (y = 4*row + rank)
    for (int i = threadIdx.x; i < 64 * 256; i += blockDim.x) {
        const int row = i >> 8, col = i & 255;
        jb.grid[(size_t)(row * 4 + rank) * 256 + col] = ((unsigned long long)hhi[i] << 32) | hlo[i];
    }
    cluster_sync_all();  // nobody exits while a peer may still signal its barriers
}

// Circular convolution for pairs with a periodic axis (convolve2D_periodic, convolve.py:215-323): the last
// row/column of a periodic axis is folded onto the first (period G-1); the reference's FFT has the size of the
// folded array in BOTH axes, so the non-periodic axis wraps as well (period G) -- reproduced here.  Rare path:
// one thread per output pixel, inputs through L2.  Same MODE semantics as k_conv2d.  grid (ceil(G*G/256), njobs).
template <int MODE>
__global__ void __launch_bounds__(256) k_conv2d_circ(const ConvJob* __restrict__ jobs, int iter) {
    const ConvJob jb = jobs[blockIdx.y];
    if (MODE == 1 && iter >= jb.mbc) return;
    if (MODE >= 2 && (!jb.lhist || (MODE == 3 && !jb.lmbc))) return;
    if (!(jb.xper | jb.yper)) return;
    const int G = jb.G, w = jb.w, K = 2 * w + 1;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    double tmax = 0;
    if (o < G * G) {
        const int y = o / G, x = o - y * G;
        const int Ny = jb.yper ? G - 1 : G, Nx = jb.xper ? G - 1 : G;
        const int yc = y % Ny, xc = x % Nx;
        const double* P = (MODE == 1) ? ((iter & 1) ? jb.Pn : jb.P) : nullptr;
        double* Pout = (MODE == 1) ? ((iter & 1) ? jb.P : jb.Pn) : jb.P;
        double thr = 0;
        if (MODE == 1) thr = __longlong_as_double((long long)jb.mx[1 + iter]) * 1e-8;
        const bool moments = (MODE == 0) && jb.bounded && jb.bco == 1;
        auto src = [&](int a, int b) {  // unfolded source element
            const double* sp = (MODE == 1) ? jb.box : (MODE == 2 ? jb.lhist : (MODE == 3 ? jb.lbox : jb.hist));
            return sp[(size_t)a * G + b];
        };
        auto folded = [&](int a, int b) {  // element (a, b) of the folded array
            double v = src(a, b);
            const bool fy = jb.yper && a == 0, fx = jb.xper && b == 0;
            if (fy) v += src(G - 1, b);
            if (fx) v += src(a, G - 1);
            if (fy && fx) v += src(G - 1, G - 1);
            return v;
        };
        double acc = 0, accx = 0, accy = 0;
        for (int ku = 0; ku < K; ku++) {
            const int u = ku - w;
            int a = (yc - u) % Ny;
            if (a < 0) a += Ny;
            for (int kv = 0; kv < K; kv++) {
                const int v = kv - w;
                int b = (xc - v) % Nx;
                if (b < 0) b += Nx;
                const double wv = jb.Wk[(size_t)ku * K + kv];
                const double c = folded(a, b);
                acc = fma(wv, c, acc);
                if (moments) {
                    accx = fma(wv * (double)v, c, accx);
                    accy = fma(wv * (double)u, c, accy);
                }
            }
        }
        if (MODE == 0) {
            jb.P[o] = acc;
            if (moments) {
                jb.xP[o] = accx;
                jb.yP[o] = accy;
            }
            tmax = acc;
        } else if (MODE == 1) {
            const double vv = P[o] * acc / jb.a00b[o];
            Pout[o] = vv;
            tmax = vv;
        } else if (MODE == 2) {
            jb.lP[o] = acc;
        } else {
            const double l1 = jb.lP[o];
            jb.lP2[o] = l1 > 0 ? acc * l1 : acc;
        }
    }
    if (MODE >= 2) return;
    tmax = warp_max(tmax);
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(jb.mx + (MODE == 0 ? 0 : 2 + iter), tmax);
}

// ----------------------------------------------------------------------------------------------------
// hot-window privatisation of the 256 x 256 histograms (default path for those grids)
// ----------------------------------------------------------------------------------------------------
// profiles/r1m: the tiled REDG kernel sits at 81 % of L2 throughput -- the L2 atomic units are the wall.  Most
// updates of a marginal 2D histogram land in a small central region (about 3/4 of a Gaussian's mass within
// +-1.5 sigma per axis = 64 x 64 bins of the 256^2 grid), so a CTA working on a 2 x 2 tile of pairs keeps the four
// 64 x 64 windows (4 x 32 KB, two 32-bit limbs per bin, native ATOMS) in shared memory, reads the pre-binned byte
// columns of k_bin8 (4 B + one weight per row for 4 pairs) and only sends the updates that fall outside a window
// to L2 as REDG.  Windows are flushed once per CTA with integer global reductions.  Exactness is unchanged:
// every path accumulates the same 64-bit fixed-point weights.
static_assert(HW == 64, "k_hist2d_hot packs dy * HW * 4 as a byte shift");
struct HotTile {
    const unsigned char* ia[2];
    const unsigned char* ib[2];
    int na, nb;
    int ax0[2], by0[2];    // window origin per parameter (column / row of the grid)
    long long off[2][2];   // grid offset of pair (a, b) or -1
};

// grid (nseg, ntiles), 256 threads, dynamic smem = 4 * HW*HW * 8 bytes
__global__ void __launch_bounds__(1024) k_hist2d_hot(const HotTile* __restrict__ tiles, const unsigned long long* __restrict__ dWq,
                                                    const Seg* __restrict__ segs, unsigned long long* __restrict__ grids) {
    extern __shared__ unsigned hsm2[];  // per window: lo[HW*HW], hi[HW*HW]
    __shared__ HotTile T;
    {
        const int* src = reinterpret_cast<const int*>(tiles + blockIdx.y);
        int* dst = reinterpret_cast<int*>(&T);
        for (int i = threadIdx.x; i < (int)(sizeof(HotTile) / 4); i += blockDim.x) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i < 4 * 2 * HW * HW; i += blockDim.x) hsm2[i] = 0;
    __syncthreads();
    const Seg sg = segs[blockIdx.x];
    const int na = T.na, nb = T.nb;
    // per-CTA constants in registers: window origins replicated into the four byte lanes, grid offsets, and the
    // 32-bit shared-memory address of window 0 (explicit .shared atomics: no generic-address arithmetic per update)
    unsigned smem0 = (unsigned)__cvta_generic_to_shared(hsm2);
    asm volatile("mov.u32 %0, %0;" : "+r"(smem0));  // opaque: keep it in a register instead of re-deriving it per update
    unsigned axr[2], byr[2];
    long long offr[2][2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
        axr[p] = (unsigned)T.ax0[p] * 0x01010101u;
        byr[p] = (unsigned)T.by0[p] * 0x01010101u;
#pragma unroll
        for (int q = 0; q < 2; q++) offr[p][q] = (p < na && q < nb) ? T.off[p][q] : -1;
    }
    auto update = [&](unsigned av, unsigned bv, int pa, int pb, unsigned long long wv) {
        const long long off = offr[pa][pb];
        if (off < 0) return;
        const unsigned dx = (av - axr[pa]) & 0xffu, dy = (bv - byr[pb]) & 0xffu;
        if ((dx | dy) < (unsigned)HW) {
            smem_add_u64_addr(smem0 + (unsigned)(pa * 2 + pb) * (2u * HW * HW * 4u) + ((dy * HW + dx) << 2), HW * HW * 4u, wv);
        } else {
            atomicAdd(grids + off + (long long)bv * 256 + av, wv);
        }
    };
    if ((sg.r0 & 3) == 0) {
        // four consecutive rows per thread: one 32-bit load per byte column, two 16-byte loads of weights
        const int64_t nq = (sg.r1 - sg.r0) >> 2;
        const bool full = offr[0][0] >= 0 && offr[0][1] >= 0 && offr[1][0] >= 0 && offr[1][1] >= 0;
        auto quads = [&](auto FULL) {
            constexpr bool kFull = decltype(FULL)::value;
            for (int64_t q = threadIdx.x; q < nq; q += blockDim.x) {
                const int64_t r = sg.r0 + 4 * q;
                unsigned a4[2] = {0, 0}, b4[2] = {0, 0};
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    if (kFull || p < na) a4[p] = *reinterpret_cast<const unsigned*>(T.ia[p] + r);
                    if (kFull || p < nb) b4[p] = *reinterpret_cast<const unsigned*>(T.ib[p] + r);
                }
                const ulonglong2 w01 = ldg_stream2_u64(dWq + r), w23 = ldg_stream2_u64(dWq + r + 2);
                // window-relative coordinates of the four rows at once (per-byte wrapping subtraction)
                unsigned da[2], db[2];
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    da[p] = __vsub4(a4[p], axr[p]);
                    db[p] = __vsub4(b4[p], byr[p]);
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned long long wv = k == 0 ? w01.x : (k == 1 ? w01.y : (k == 2 ? w23.x : w23.y));
                    if (wv == 0) continue;
#pragma unroll
                    for (int pa = 0; pa < 2; pa++)
#pragma unroll
                        for (int pb = 0; pb < 2; pb++) {
                            if (!kFull && offr[pa][pb] < 0) continue;  // CTA-uniform
                            if (((da[pa] | db[pb]) & (0xC0u << (8 * k))) == 0) {  // both coordinates < 64: inside the window
                                const unsigned dx = __byte_perm(da[pa], 0, 0x4440 | k);          // byte k -> bits 0..7
                                const unsigned dy8 = __byte_perm(db[pb], 0, 0x4404 | (k << 4));  // byte k -> bits 8..15 = dy * HW * 4
                                smem_add_u64_addr(smem0 + (unsigned)(pa * 2 + pb) * (2u * HW * HW * 4u) + dy8 + (dx << 2), HW * HW * 4u, wv);
                            } else {
                                const unsigned av = (a4[pa] >> (8 * k)) & 0xffu, bv = (b4[pb] >> (8 * k)) & 0xffu;
                                atomicAdd(grids + offr[pa][pb] + (long long)bv * 256 + av, wv);
                            }
                        }
                }
            }
        };
        if (full)
            quads(std::true_type{});
        else
            quads(std::false_type{});
        for (int64_t r = sg.r0 + 4 * nq + threadIdx.x; r < sg.r1; r += blockDim.x) {  // tail rows
            const unsigned long long wv = dWq[r];
            if (wv == 0) continue;
            for (int pa = 0; pa < na; pa++)
                for (int pb = 0; pb < nb; pb++) update(T.ia[pa][r], T.ib[pb][r], pa, pb, wv);
        }
    } else {
        for (int64_t r = sg.r0 + threadIdx.x; r < sg.r1; r += blockDim.x) {
            const unsigned long long wv = dWq[r];
            if (wv == 0) continue;
            for (int pa = 0; pa < na; pa++)
                for (int pb = 0; pb < nb; pb++) update(T.ia[pa][r], T.ib[pb][r], pa, pb, wv);
        }
    }
    __syncthreads();
    for (int q = 0; q < 4; q++) {
        const int pa = q >> 1, pb = q & 1;
        if (pa >= na || pb >= nb) continue;
        const long long off = T.off[pa][pb];
        if (off < 0) continue;
        const unsigned* base = hsm2 + q * 2 * HW * HW;
        for (int i = threadIdx.x; i < HW * HW; i += blockDim.x) {
            const unsigned long long v = ((unsigned long long)base[HW * HW + i] << 32) | base[i];
            if (v) atomicAdd(grids + off + (long long)(T.by0[pb] + (i / HW)) * 256 + T.ax0[pa] + (i % HW), v);
        }
    }
}

// box = hist / P where P > 1e-8 max(P), else hist (mcsamples.py:1969-1971): the input of bias iteration `iter`.
// grid (64, njobs)
__global__ void __launch_bounds__(256) k_make_box(const ConvJob* __restrict__ jobs, int iter) {
    const ConvJob jb = jobs[blockIdx.y];
    if (iter >= jb.mbc) return;
    const size_t gg = (size_t)jb.G * jb.G;
    const double* P = (iter & 1) ? jb.Pn : jb.P;
    const double thr = __longlong_as_double((long long)jb.mx[1 + iter]) * 1e-8;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < gg; o += (size_t)gridDim.x * blockDim.x) {
        const double h = jb.hist[o], p = P[o];
        jb.box[o] = p > thr ? h / p : h;
    }
}

// ----------------------------------------------------------------------------------------------------
// bucket-sorted sweep for the 256 x 256 histograms (default path for those grids)
// ----------------------------------------------------------------------------------------------------
// profiles/r1s: k_hist2d_hot is bound by shared-memory atomics (L1 85 %, ~3.7 bank-conflict wavefronts per
// update because the 32 lanes of a warp hit random bins) plus the 27 % of updates that miss the windows and go
// to L2.  Here the rows are counting-sorted by the bin of an ANCHOR parameter a (256 buckets), so all rows of
// bucket c update only column c (or row c) of the grids (a, b): 256 bins per partner b.  A CTA keeps those 256
// bins for 32 lanes in shared memory laid out [bin][lane]: the lane index is the bank, so every shared atomic is
// conflict-free and two lanes of a warp never share an address.  A lane is one (row-slot, partner) combination:
// with nl <= 16 partners a warp processes 32 / nlp rows per step into per-lane replica histograms that the flush
// sums.  Every pair (i, j) is covered once by the circular rule "anchor i has partners i+1 .. i+P/2 (mod P)".
//
// The sort is MATERIALISED (profiles/r2a: gathering rows through a permutation cost 196 B of DRAM traffic per row
// visit, 125 GB per triangle): k_bucket_records writes, for every (anchor, 32 partners) job, one 32-byte record
// per row -- the partners' byte bins in lane order -- plus the 64-bit fixed-point weight at the row's position in
// bucket order; k_hist2d_records then streams records and weights sequentially (40 B per row visit).
// Buckets with few rows in a chunk (distribution tails) go straight to L2 reductions.  Same fixed-point weights
// as every other path: bit-identical histograms.
#define SRT_THREADS 512
#define SRT_BIG 96      // rows of one bucket inside a chunk from which the shared-memory path pays for its flush
#define SRT_MAXJOBS 32  // jobs per batch (shared-memory tables of k_bucket_records: 32 x 256 x 8 B)
struct SortJob {
    int slot, nl, lg, c0;  // anchor slot, partners (<= 32), log2(lanes per row), first partner column if contiguous else -1
    int pcol[32];          // partner column (slot) per lane
    int sb[32], sc[32];    // grid strides of the partner's bin / of the anchor's bucket
    long long off[32];     // grid offsets (elements)
};

// k_bin8 + per-parameter bucket counts.  grid (nseg, nparams), 256 threads.
__global__ void __launch_bounds__(256) k_bin8c(const double* __restrict__ dX, int64_t ld, const Seg* __restrict__ segs,
                                               const Bin8Job* __restrict__ jobs, unsigned char* __restrict__ out, int64_t old,
                                               unsigned* __restrict__ counts) {
    __shared__ unsigned cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const Bin8Job jb = jobs[blockIdx.y];
    const Seg sg = segs[blockIdx.x];
    const double* x = dX + (int64_t)jb.param * ld;
    unsigned char* o = out + (int64_t)blockIdx.y * old;
    for (int64_t r0 = sg.r0 + threadIdx.x; r0 < sg.r1; r0 += 4 * (int64_t)blockDim.x) {
        double xv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {  // four independent loads in flight
            const int64_t r = r0 + (int64_t)k * blockDim.x;
            xv[k] = r < sg.r1 ? ldg_stream(x + r) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t r = r0 + (int64_t)k * blockDim.x;
            if (r >= sg.r1) break;
            const int b = bin_index_round(xv[k], jb.binmin, jb.fw, jb.inv);
            const int c = b < 0 ? 0 : (b > 255 ? 255 : b);
            o[r] = (unsigned char)c;
            atomicAdd(&cnt[c], 1u);
        }
    }
    __syncthreads();
    if (cnt[threadIdx.x]) atomicAdd(counts + (size_t)blockIdx.y * 256 + threadIdx.x, cnt[threadIdx.x]);
}

// exclusive scan of the 256 bucket counts of each parameter.  grid nparams, 256 threads.
// start[p*257 + c] = first position of bucket c (start[p*257 + 256] = N).
__global__ void __launch_bounds__(256) k_bucket_scan(const unsigned* __restrict__ counts, unsigned* __restrict__ start) {
    __shared__ unsigned s[256];
    const int t = threadIdx.x, p = blockIdx.x;
    const unsigned own = counts[(size_t)p * 256 + t];
    s[t] = own;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        const unsigned v = t >= d ? s[t - d] : 0u;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    start[(size_t)p * 257 + t] = s[t] - own;
    if (t == 255) start[(size_t)p * 257 + 256] = s[255];
}

// write cursors of a batch of jobs: cursor[j*256 + c] = start[slot_j*257 + c].  grid njobs, 256 threads.
__global__ void __launch_bounds__(256) k_cursor_init(const SortJob* __restrict__ jobs, const unsigned* __restrict__ start,
                                                     unsigned* __restrict__ cursor) {
    cursor[(size_t)blockIdx.x * 256 + threadIdx.x] = start[(size_t)jobs[blockIdx.x].slot * 257 + threadIdx.x];
}

// counting-sort scatter for a batch of jobs.  grid ceil(N / rows), 1024 threads.  A CTA stages its rows of the
// column-major byte bins as a row-major tile in shared memory (row pitch `pitch`: np + 36 bytes so that the 32
// partner columns of a contiguous circular window never wrap; pitch / 4 is odd -> conflict-free row-strided
// loads), counts the rows per (job, bucket), reserves a range of every non-empty bucket with one global atomic,
// and then, job by job, sorts its rows by bucket INSIDE shared memory and copies the sorted run out: the records
// of one bucket land next to each other, so consecutive lanes write consecutive 16-byte halves (profiles/r2b: one
// scattered 32-byte record + weight per thread ran at 1.3 TB/s of DRAM writes, bound by store-sector requests).
// Order inside a bucket is arbitrary (integer sums).
// dynamic smem: rows * pitch (tile) + njobs * 256 * 8 (counts -> global-minus-local offsets, local starts)
//               + rows * 44 (staged records, weights, positions) + 2 * 256 * 4 (local cursors)
__global__ void __launch_bounds__(1024) k_bucket_records(const unsigned char* __restrict__ ix8, int64_t ld, int np, int pitch,
                                                         int rows, int64_t N, const unsigned long long* __restrict__ wq,
                                                         const SortJob* __restrict__ jobs, int njobs, unsigned* __restrict__ cursor,
                                                         uint4* __restrict__ recs, unsigned long long* __restrict__ ws, int64_t pld) {
    extern __shared__ __align__(16) unsigned char rsm[];
    uint4* stage_rec = reinterpret_cast<uint4*>(rsm);                                            // [rows][2]
    unsigned long long* stage_w = reinterpret_cast<unsigned long long*>(rsm + (size_t)rows * 32);  // [rows]
    unsigned* stage_pos = reinterpret_cast<unsigned*>(rsm + (size_t)rows * 40);                  // [rows]
    unsigned* lcur = reinterpret_cast<unsigned*>(rsm + (size_t)rows * 44);                       // [2][256]
    unsigned* gdel = lcur + 512;                                                                  // [njobs][256]
    unsigned* lstart = gdel + (size_t)njobs * 256;                                                // [njobs][256]
    unsigned char* tile = reinterpret_cast<unsigned char*>(lstart + (size_t)njobs * 256);         // [rows][pitch]
    __shared__ int jslot[SRT_MAXJOBS], jc0[SRT_MAXJOBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * rows;
    const int nrow = (int)min((int64_t)rows, N - r0);
    for (int i = threadIdx.x; i < njobs * 256; i += blockDim.x) gdel[i] = 0;
    if (threadIdx.x < njobs) {
        jslot[threadIdx.x] = jobs[threadIdx.x].slot;
        jc0[threadIdx.x] = jobs[threadIdx.x].c0;
    }
    // stage.  np % 4 == 0: a warp item = 16 parameters x 32 rows; a lane loads the words of 4 parameters x 4 rows,
    // transposes the 4 x 4 bytes in registers and stores one word per row (bank = 4 ql (pitch/4 mod 8) + pl: all 32
    // lanes distinct because pitch / 4 is odd).  Otherwise (small odd P): byte stores.
    if ((np & 3) == 0) {
        const int ql = lane & 7, pl = lane >> 3;
        const int npg = (np + 15) >> 4, nqb = (rows + 31) >> 5;
        for (int it = warp; it < npg * nqb; it += nwarps) {
            const int pg = it / nqb, qb = it - pg * nqb;
            const int p0 = pg * 16 + pl * 4, q = qb * 8 + ql;  // rows 4q .. 4q+3
            if (p0 < np && 4 * q < rows) {
                unsigned wv[4];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    wv[j] = (r0 + 4 * q < ld) ? *reinterpret_cast<const unsigned*>(ix8 + (int64_t)(p0 + j) * ld + r0 + 4 * q) : 0u;
                const unsigned t0 = __byte_perm(wv[0], wv[1], 0x5140), t1 = __byte_perm(wv[0], wv[1], 0x7362);
                const unsigned t2 = __byte_perm(wv[2], wv[3], 0x5140), t3 = __byte_perm(wv[2], wv[3], 0x7362);
                const unsigned o[4] = {__byte_perm(t0, t2, 0x5410), __byte_perm(t0, t2, 0x7632), __byte_perm(t1, t3, 0x5410),
                                       __byte_perm(t1, t3, 0x7632)};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    unsigned* row = reinterpret_cast<unsigned*>(tile + (size_t)(4 * q + k) * pitch);
                    row[p0 >> 2] = o[k];
                    if (p0 < 32) row[(np + p0) >> 2] = o[k];  // wrap-around copy of the first 32 columns
                }
            }
        }
    } else {
        const int ql = lane & 7, pl = lane >> 3;
        const int npg = (np + 3) >> 2, nqb = (rows + 31) >> 5;
        for (int it = warp; it < npg * nqb; it += nwarps) {
            const int pg = it / nqb, qb = it - pg * nqb;
            const int p = pg * 4 + pl, q = qb * 8 + ql;  // word q = rows 4q .. 4q+3 of the tile
            if (p < np && 4 * q < rows) {
                const unsigned wv = (r0 + 4 * q < ld) ? *reinterpret_cast<const unsigned*>(ix8 + (int64_t)p * ld + r0 + 4 * q) : 0u;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    unsigned char* row = tile + (size_t)(4 * q + k) * pitch;
                    const unsigned char b = (unsigned char)(wv >> (8 * k));
                    row[p] = b;
                    if (p < 32) row[np + p] = b;  // wrap-around copy of the first 32 columns
                }
            }
        }
    }
    __syncthreads();
    // phase A: rows per (job, bucket)
    for (int t = threadIdx.x; t < nrow; t += blockDim.x) {
        const unsigned char* row = tile + (size_t)t * pitch;
        for (int j = 0; j < njobs; j++) atomicAdd(&gdel[j * 256 + row[jslot[j]]], 1u);
    }
    __syncthreads();
    // phase B: one warp per job: exclusive scan of the 256 counts (8 buckets per lane) -> local starts; one global
    // atomic per non-empty bucket reserves [base, base + count); gdel = base - local start
    for (int j = warp; j < njobs; j += nwarps) {
        unsigned v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[k] = gdel[j * 256 + lane * 8 + k];
            sum += v[k];
        }
        unsigned incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        unsigned pre = incl - sum;
        unsigned base[8];
#pragma unroll
        for (int k = 0; k < 8; k++) base[k] = v[k] ? atomicAdd(cursor + (size_t)j * 256 + lane * 8 + k, v[k]) : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            lstart[j * 256 + lane * 8 + k] = pre;
            gdel[j * 256 + lane * 8 + k] = base[k] - pre;
            pre += v[k];
        }
    }
    __syncthreads();
    // phase C, job by job: local counting sort into the staging buffers, then a coalesced copy-out (rows <= 1024:
    // one row per thread)
    const bool has_row = (int)threadIdx.x < nrow;
    const unsigned char* row = tile + (size_t)threadIdx.x * pitch;
    const unsigned long long wrow_q = has_row ? wq[r0 + threadIdx.x] : 0ull;
    if (threadIdx.x < 256 && njobs > 0) lcur[threadIdx.x] = lstart[threadIdx.x];
    for (int j = 0; j < njobs; j++) {
        unsigned* cur = lcur + (j & 1) * 256;
        __syncthreads();  // cursors of job j ready; copy-out of job j-1 has finished reading the staging buffers
        const int slot = jslot[j], c0 = jc0[j];
        if (has_row) {
            const unsigned c = row[slot];
            const unsigned li = atomicAdd(&cur[c], 1u);
            unsigned v[8];
            if (c0 >= 0) {  // contiguous window [c0, c0 + 32): nine aligned words, funnel-shifted
                const unsigned* wrow = reinterpret_cast<const unsigned*>(row) + (c0 >> 2);
                const unsigned sh = (unsigned)(c0 & 3) * 8u;
                unsigned prev = wrow[0];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const unsigned next = wrow[k + 1];
                    v[k] = __funnelshift_r(prev, next, sh);
                    prev = next;
                }
            } else {
                const int* pc = jobs[j].pcol;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    v[k] = (unsigned)row[pc[4 * k]] | ((unsigned)row[pc[4 * k + 1]] << 8) | ((unsigned)row[pc[4 * k + 2]] << 16) |
                           ((unsigned)row[pc[4 * k + 3]] << 24);
            }
            stage_rec[2 * li] = make_uint4(v[0], v[1], v[2], v[3]);
            stage_rec[2 * li + 1] = make_uint4(v[4], v[5], v[6], v[7]);
            stage_w[li] = wrow_q;
            stage_pos[li] = li + gdel[j * 256 + c];
        }
        __syncthreads();
        uint4* dstj = recs + (size_t)j * pld * 2;
        for (int i = threadIdx.x; i < 2 * nrow; i += blockDim.x) dstj[(size_t)stage_pos[i >> 1] * 2 + (i & 1)] = stage_rec[i];
        for (int i = threadIdx.x; i < nrow; i += blockDim.x) ws[(size_t)j * pld + stage_pos[i]] = stage_w[i];
        if (threadIdx.x < 256 && j + 1 < njobs) lcur[((j + 1) & 1) * 256 + threadIdx.x] = lstart[(j + 1) * 256 + threadIdx.x];
    }
}

// grid (nchunks, njobs), SRT_THREADS threads, dynamic smem = 2 * 256 * 32 * 4 bytes (lo limbs, hi limbs; [bin][lane]).
// LG = log2(lanes per row) of every job of the launch (the host groups jobs by it).
template <int LG>
__global__ void __launch_bounds__(SRT_THREADS, 2)
    k_hist2d_records(const SortJob* __restrict__ jobs, int job0, const unsigned char* __restrict__ recs,
                     const unsigned long long* __restrict__ ws, int64_t pld, const unsigned* __restrict__ start,
                     unsigned long long* __restrict__ grids, int chunk, int64_t N) {
    extern __shared__ unsigned ssrt[];  // lo[256 * 32], hi[256 * 32]
    __shared__ SortJob J;
    __shared__ unsigned sst[257];
    const int jix = job0 + blockIdx.y;  // job index inside the batch = index of its record block
    {
        const int* src = reinterpret_cast<const int*>(jobs + jix);
        int* dst = reinterpret_cast<int*>(&J);
        for (int i = threadIdx.x; i < (int)(sizeof(SortJob) / 4); i += blockDim.x) dst[i] = src[i];
    }
    for (int i = threadIdx.x; i < 2 * 256 * 32; i += blockDim.x) ssrt[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < 257; i += blockDim.x) sst[i] = start[(size_t)J.slot * 257 + i];
    __syncthreads();
    constexpr int NLP = 1 << LG, R = 32 >> LG;                  // lanes per row, rows per step
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int rs = lane >> LG, pl = lane & (NLP - 1);          // row slot of this lane, partner
    const bool act = pl < J.nl;
    const long long off = act ? J.off[pl] : 0;
    const int sb = act ? J.sb[pl] : 0, sc = act ? J.sc[pl] : 0;
    const unsigned char* rj = recs + ((size_t)jix * pld + rs) * 32 + pl;  // + position * 32: this lane's partner byte of row slot rs
    const unsigned long long* wj = ws + (size_t)jix * pld + rs;
    unsigned smem_lane = (unsigned)__cvta_generic_to_shared(ssrt) + (unsigned)lane * 4u;
    asm volatile("mov.u32 %0, %0;" : "+r"(smem_lane));
    const int64_t cs = (int64_t)blockIdx.x * chunk, ce = min(N, cs + (int64_t)chunk);

    // positions [lo, hi) of bucket c into the shared bins: blocks of 32 positions per warp, NLP steps of R rows
    auto rows_smem = [&](int64_t lo, int64_t hi) {
        for (int64_t b0 = lo + (int64_t)warp * 32; b0 < hi; b0 += (int64_t)nwarps * 32) {
            const unsigned char* rb = rj + (size_t)b0 * 32;
            const unsigned long long* wb = wj + b0;
            if (b0 + 32 <= hi) {
                if (!act) continue;
#pragma unroll
                for (int it0 = 0; it0 < NLP; it0 += 4) {  // four steps of loads in flight
                    unsigned bv[4];
                    unsigned long long wv[4];
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (it0 + u < NLP) {
                            bv[u] = __ldg(rb + (it0 + u) * R * 32);
                            wv[u] = __ldg(wb + (it0 + u) * R);  // same address across the lanes of a row: broadcast
                        }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (it0 + u < NLP) smem_add_u64_addr(smem_lane + bv[u] * 128u, 256u * 32u * 4u, wv[u]);
                }
            } else if (act) {
                for (int it = 0; it < NLP; it++)
                    if (b0 + it * R + rs < hi)
                        smem_add_u64_addr(smem_lane + (unsigned)rb[it * R * 32] * 128u, 256u * 32u * 4u, wb[it * R]);
            }
        }
    };
    // positions [lo, hi) of a run of small buckets starting with bucket c: straight to the grids in L2
    auto rows_global = [&](int64_t lo, int64_t hi, int c) {
        if (!act) return;
        for (int64_t b0 = lo + (int64_t)warp * 32; b0 < hi; b0 += (int64_t)nwarps * 32)
            for (int it = 0; it < NLP; it++) {
                const int64_t pos = b0 + it * R + rs;
                if (pos >= hi) continue;
                const unsigned long long w = wj[pos - rs];
                if (w == 0) continue;
                int cc = c;  // bucket of the position: the last c' with sst[c'] <= position
                while (cc < 255 && (int64_t)sst[cc + 1] <= pos) cc++;
                atomicAdd(grids + off + (long long)rj[(size_t)(pos - rs) * 32] * sb + (long long)cc * sc, w);
            }
    };
    auto flush = [&](int c) {
        for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {  // i & 31 == lane: this thread's own partner
            const unsigned vlo = ssrt[i], vhi = ssrt[256 * 32 + i];
            if (vlo | vhi) {
                atomicAdd(grids + off + (long long)(i >> 5) * sb + (long long)c * sc, ((unsigned long long)vhi << 32) | vlo);
                ssrt[i] = 0;
                ssrt[256 * 32 + i] = 0;
            }
        }
    };
    // first bucket that reaches into this chunk: largest c with sst[c] <= cs
    int c = 0;
    {
        int lo = 0, hi = 256;  // invariant: sst[lo] <= cs, (hi == 256 or sst[hi] > cs)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((int64_t)sst[mid] <= cs) lo = mid; else hi = mid;
        }
        c = lo;
    }
    while (c < 256) {
        const int64_t s0 = sst[c], s1 = sst[c + 1];
        if (s0 >= ce) break;
        const int64_t lo = max(s0, cs), hi = min(s1, ce);
        if (hi <= lo) {
            c++;
            continue;
        }
        if (hi - lo >= SRT_BIG) {
            rows_smem(lo, hi);
            __syncthreads();
            flush(c);
            __syncthreads();
            c++;
        } else {  // a run of small buckets: one pass, straight to L2
            int c2 = c + 1;
            int64_t end = hi;
            while (c2 < 256 && (int64_t)sst[c2] < ce) {
                const int64_t e2 = min((int64_t)sst[c2 + 1], ce);
                if (e2 - (int64_t)sst[c2] >= SRT_BIG) break;
                end = e2;
                c2++;
            }
            rows_global(lo, end, c);
            c = c2;
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// bucket-sorted sweep for the sheared re-binning (kde.bin_samples of (p1, r0*x_i + r1*x_j), 256 x 256 grids)
// ----------------------------------------------------------------------------------------------------
// profiles/r1s: k_shear_minmax streams 52 GB (every job re-reads both of its columns; DRAM bound, 9.5 ms) and
// k_shear_hist another 63 GB with hot-window atomics (21.5 ms).  Here anchors (= p1 columns) whose partner sets
// overlap are batched: a CTA stages a tile of rows x (<= SHR_MAXCOLS distinct columns) float64 values in shared
// memory ONCE per pass and evaluates every sheared pair of the batch from it (11 GB per pass at C2).
//   pass 1 (k_shear_minmax_tiled): min / max of p2 per job + the bucket counts of b1 per anchor;
//   pass 2 (k_shear_records): b1 -> bucket, the partners' b2 bins -> one 32-byte record per (row, anchor), written
//          in bucket order exactly like k_bucket_records; k_hist2d_records then sweeps the records.
// p2 is evaluated with the same explicit multiply / multiply / add as everywhere else (no FMA contraction).
#define SHR_ROWS 512
#define SHR_MAXCOLS 18
#define SHR_MAXJOBS 16
struct ShearPairRef {  // one sheared job inside a batch
    int acol, pcol, job, anchor;  // tile columns of x_i and x_j, global job index, anchor index inside the batch
    double r0, r1;
};
struct ShearRecJob {  // one anchor (p1 column + geometry) with <= 32 partners
    int acol, np, pair0, pad;  // tile column of p1, partners, first ShearPairRef of this anchor (consecutive)
    double p1_min, dx1, inv1s;  // inv1s = 2^20 / dx1
};
#define SHR_MAXUNITS 32  // k_shear_minmax_tma: 16 warps x 2 units of <= 4 pairs that share their p1 column
struct ShearBatch {
    int ncols, njobs, npairs, job0;  // job0: index of the batch's first anchor (row of counts / start / cursor)
    int pair0, nunits, pad[2];        // first ShearPairRef of the batch; units of k_shear_minmax_tma
    int cols[SHR_MAXCOLS];            // parameter (column of dX) per tile column
    short ufirst[SHR_MAXUNITS];       // unit -> first pair (index inside the batch), its pairs are consecutive
    unsigned char unp[SHR_MAXUNITS];  // unit -> number of pairs (1..4)
    unsigned char uacol[SHR_MAXUNITS];  // unit -> tile column of p1
};

// grid (nseg, nbatches), 512 threads.  dynamic smem: ncols * SHR_ROWS * 8 (tile) + npairs * 32 (pair table) +
// njobs * 1024 (b1 bucket counts, only when counts != NULL).
// part[(job * nseg + seg) * 2 + {0, 1}] = min, max of p2 over the segment.  Pair i of the batch belongs to warp
// i % 16, which keeps its running min / max in registers across all tiles of the segment (<= SHR_MAXPW pairs per
// warp) and reduces across lanes once at the end.
#define SHR_MAXPW 8  // pairs per warp: batches hold <= 16 * SHR_MAXPW pairs
__global__ void __launch_bounds__(512) k_shear_minmax_tiled(const double* __restrict__ dX, int64_t ld, const Seg* __restrict__ segs,
                                                            int nseg, const ShearBatch* __restrict__ batches,
                                                            const ShearRecJob* __restrict__ jobs, const ShearPairRef* __restrict__ pairs,
                                                            double* __restrict__ part, unsigned* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char msm[];
    __shared__ ShearBatch B;
    {
        const int* src = reinterpret_cast<const int*>(batches + blockIdx.y);
        int* dst = reinterpret_cast<int*>(&B);
        for (int i = threadIdx.x; i < (int)(sizeof(ShearBatch) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    double* tile = reinterpret_cast<double*>(msm);                                      // [ncols][SHR_ROWS]
    ShearPairRef* sp = reinterpret_cast<ShearPairRef*>(tile + (size_t)B.ncols * SHR_ROWS);  // [npairs]
    unsigned* cnt = reinterpret_cast<unsigned*>(sp + B.npairs);                         // [njobs][256]
    const ShearRecJob* bj = jobs + B.job0;
    for (int i = threadIdx.x; i < B.npairs; i += blockDim.x) sp[i] = pairs[B.pair0 + i];
    if (counts)
        for (int i = threadIdx.x; i < B.njobs * 256; i += blockDim.x) cnt[i] = 0;
    const Seg sg = segs[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double mn[SHR_MAXPW], mx[SHR_MAXPW];
#pragma unroll
    for (int k = 0; k < SHR_MAXPW; k++) {
        mn[k] = INFINITY;
        mx[k] = -INFINITY;
    }
    for (int64_t t0 = sg.r0; t0 < sg.r1; t0 += SHR_ROWS) {
        const int nrow = (int)min((int64_t)SHR_ROWS, sg.r1 - t0);
        __syncthreads();  // previous tile fully consumed (and the tables above are visible)
        for (int c = 0; c < B.ncols; c++) {
            const double* col = dX + (int64_t)B.cols[c] * ld + t0;
            for (int t = threadIdx.x; t < nrow; t += blockDim.x) tile[c * SHR_ROWS + t] = ldg_stream(col + t);
        }
        __syncthreads();
        if (counts)  // bucket counts of b1 per anchor (one row per thread)
            for (int t = threadIdx.x; t < nrow; t += blockDim.x)
                for (int j = 0; j < B.njobs; j++) {
                    const int b1 = bin_index_trunc_fx(tile[bj[j].acol * SHR_ROWS + t], bj[j].p1_min, bj[j].dx1, bj[j].inv1s);
                    atomicAdd(&cnt[j * 256 + min(max(b1, 0), 255)], 1u);
                }
#pragma unroll
        for (int k = 0; k < SHR_MAXPW; k++) {
            const int i = warp + k * 16;
            if (i < B.npairs) {
                const ShearPairRef pr = sp[i];
                const double* xi = tile + pr.acol * SHR_ROWS + lane;
                const double* xj = tile + pr.pcol * SHR_ROWS + lane;
                if (nrow == SHR_ROWS) {
#pragma unroll
                    for (int m = 0; m < SHR_ROWS / 32; m++) {
                        const double p2 = shear_p2(xi[32 * m], xj[32 * m], pr.r0, pr.r1);
                        mn[k] = p2 < mn[k] ? p2 : mn[k];  // finite samples: plain compares, not the NaN-aware fmin / fmax
                        mx[k] = p2 > mx[k] ? p2 : mx[k];
                    }
                } else {
                    for (int t = lane; t < nrow; t += 32) {
                        const double p2 = shear_p2(xi[t - lane], xj[t - lane], pr.r0, pr.r1);
                        mn[k] = p2 < mn[k] ? p2 : mn[k];
                        mx[k] = p2 > mx[k] ? p2 : mx[k];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < SHR_MAXPW; k++) {
        const int i = warp + k * 16;
        if (i < B.npairs) {
            const double a = warp_min(mn[k]), b = warp_max(mx[k]);
            if (lane == 0) {
                part[((int64_t)sp[i].job * nseg + blockIdx.x) * 2 + 0] = a;
                part[((int64_t)sp[i].job * nseg + blockIdx.x) * 2 + 1] = b;
            }
        }
    }
    if (counts) {
        __syncthreads();
        for (int i = threadIdx.x; i < B.njobs * 256; i += blockDim.x)
            if (cnt[i]) atomicAdd(counts + (size_t)B.job0 * 256 + i, cnt[i]);
    }
}

// TMA-pipelined form of the min / max pass (default when every segment starts on an even row): the row tiles
// (<= SHR_MAXCOLS columns x SHR_ROWS rows of float64) are staged through a SHM_STAGES-deep shared-memory ring by bulk
// async copies (cp.async.bulk -> UBLKCP) with full / empty mbarriers, so the next tile lands while the current one is
// evaluated.  A warp owns two UNITS = (p1 column, <= 4 partners): x_i is read once per row and unit, the partners'
// r0 / r1 and the running min / max stay in registers across all tiles of the segment (11 instructions per
// pair-sample, 5 of them on the FP64 pipe, which is what bounds the pass once the loads are off the critical path).
// grid (nseg, nbatches), 544 threads (16 consumer warps + 1 producer warp), one CTA per SM;
// dynamic smem: SHM_STAGES * ncols * SHR_ROWS * 8 + npairs * 32.
#define SHM_STAGES 2
template <int NPU>
__device__ __forceinline__ void shear_mm_unit(const double* __restrict__ tl, const ShearPairRef* __restrict__ pr, int acol, int lane,
                                              int nrow, double (&mn)[4], double (&mx)[4]) {
    double r0[NPU], r1[NPU];
    const double* xj[NPU];
#pragma unroll
    for (int k = 0; k < NPU; k++) {
        r0[k] = pr[k].r0;
        r1[k] = pr[k].r1;
        xj[k] = tl + pr[k].pcol * SHR_ROWS + lane;
    }
    const double* xi = tl + acol * SHR_ROWS + lane;
    if (nrow == SHR_ROWS) {
#pragma unroll
        for (int m = 0; m < SHR_ROWS / 32; m++) {
            const double a = xi[32 * m];
#pragma unroll
            for (int k = 0; k < NPU; k++) {
                const double p2 = shear_p2(a, xj[k][32 * m], r0[k], r1[k]);
                mn[k] = p2 < mn[k] ? p2 : mn[k];  // finite samples: plain compares, not the NaN-aware fmin / fmax
                mx[k] = p2 > mx[k] ? p2 : mx[k];
            }
        }
    } else {
        for (int t = lane; t < nrow; t += 32) {
            const double a = xi[t - lane];
#pragma unroll
            for (int k = 0; k < NPU; k++) {
                const double p2 = shear_p2(a, xj[k][t - lane], r0[k], r1[k]);
                mn[k] = p2 < mn[k] ? p2 : mn[k];
                mx[k] = p2 > mx[k] ? p2 : mx[k];
            }
        }
    }
}

__global__ void __launch_bounds__(544, 1) k_shear_minmax_tma(const double* __restrict__ dX, int64_t ld, const Seg* __restrict__ segs,
                                                             int nseg, const ShearBatch* __restrict__ batches,
                                                             const ShearPairRef* __restrict__ pairs, double* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char msm[];
    __shared__ ShearBatch B;
    __shared__ __align__(8) unsigned long long full[SHM_STAGES], empty[SHM_STAGES];
    {
        const int* src = reinterpret_cast<const int*>(batches + blockIdx.y);
        int* dst = reinterpret_cast<int*>(&B);
        for (int i = threadIdx.x; i < (int)(sizeof(ShearBatch) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int ncols = B.ncols;
    double* tile = reinterpret_cast<double*>(msm);                                                      // [stage][ncols][SHR_ROWS]
    ShearPairRef* sp = reinterpret_cast<ShearPairRef*>(tile + (size_t)SHM_STAGES * ncols * SHR_ROWS);   // [npairs]
    for (int i = threadIdx.x; i < B.npairs; i += blockDim.x) sp[i] = pairs[B.pair0 + i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < SHM_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 16);  // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const Seg sg = segs[blockIdx.x];
    const int64_t n = sg.r1 - sg.r0;
    const int ntiles = (int)((n + SHR_ROWS - 1) / SHR_ROWS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto issue = [&](int c) {  // executed by the producer warp: lane l copies column l, l + 32, ...
        const int s = c % SHM_STAGES;
        const int64_t off = (int64_t)c * SHR_ROWS;
        const int cnt = (int)min((int64_t)SHR_ROWS, n - off);
        const unsigned bytes = (unsigned)(((cnt + 1) & ~1) * 8);  // columns are padded: one element past is readable
        if (lane == 0) mbar_expect_tx(&full[s], bytes * (unsigned)ncols);
        __syncwarp();
        for (int col = lane; col < ncols; col += 32)
            bulk_g2s(tile + ((size_t)s * ncols + col) * SHR_ROWS, dX + (int64_t)B.cols[col] * ld + sg.r0 + off, bytes, &full[s]);
    };
    if (warp == 16) {  // producer warp: keeps the ring full, never computes
        for (int c = 0; c < ntiles; c++) {
            if (c >= SHM_STAGES) mbar_wait(&empty[c % SHM_STAGES], (unsigned)(((c / SHM_STAGES) - 1) & 1));
            issue(c);
        }
        return;
    }
    double mn[2][4], mx[2][4];
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            mn[u][k] = INFINITY;
            mx[u][k] = -INFINITY;
        }
    for (int c = 0; c < ntiles; c++) {
        const int s = c % SHM_STAGES;
        const unsigned ph = (unsigned)((c / SHM_STAGES) & 1);
        mbar_wait(&full[s], ph);
        const int nrow = (int)min((int64_t)SHR_ROWS, n - (int64_t)c * SHR_ROWS);
        const double* tl = tile + (size_t)s * ncols * SHR_ROWS;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int unit = warp * 2 + u;
            if (unit < B.nunits) {
                const ShearPairRef* pr = sp + B.ufirst[unit];
                const int acol = B.uacol[unit];
                switch (B.unp[unit]) {
                    case 1: shear_mm_unit<1>(tl, pr, acol, lane, nrow, mn[u], mx[u]); break;
                    case 2: shear_mm_unit<2>(tl, pr, acol, lane, nrow, mn[u], mx[u]); break;
                    case 3: shear_mm_unit<3>(tl, pr, acol, lane, nrow, mn[u], mx[u]); break;
                    default: shear_mm_unit<4>(tl, pr, acol, lane, nrow, mn[u], mx[u]); break;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the stage
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int unit = warp * 2 + u;
        if (unit < B.nunits) {
            const int np = B.unp[unit], first = B.ufirst[unit];
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < np) {
                    const double a = warp_min(mn[u][k]), b = warp_max(mx[u][k]);
                    if (lane == 0) {
                        part[((int64_t)sp[first + k].job * nseg + blockIdx.x) * 2 + 0] = a;
                        part[((int64_t)sp[first + k].job * nseg + blockIdx.x) * 2 + 1] = b;
                    }
                }
        }
    }
}

// grid ceil(N / SHR_ROWS), 512 threads, one launch per batch.  Structure of k_bucket_records (count, reserve, local
// sort, coalesced copy-out) with the record bytes computed from the float64 tile.  Rows whose b1 falls outside the
// grid keep weight 0 (kde.bin_samples ranges cover the samples, so this only guards bad hard limits).
// dynamic smem: SHR_ROWS * 44 + 2048 + njobs * (2048 + SHR_ROWS) + 32 * 48 + ncols * SHR_ROWS * 8
struct ShearLaneParam {
    double r0, r1, rmin, dx, invs;  // invs = 2^20 / dx of the job's p2 range
    int pcol, pad;
};
__global__ void __launch_bounds__(512) k_shear_records(const double* __restrict__ dX, int64_t ld, int64_t N,
                                                       const unsigned long long* __restrict__ wq, const ShearBatch* __restrict__ batches,
                                                       int batch, const ShearRecJob* __restrict__ jobs,
                                                       const ShearPairRef* __restrict__ pairs, const ShearGeom* __restrict__ geom,
                                                       unsigned* __restrict__ cursor, uint4* __restrict__ recs,
                                                       unsigned long long* __restrict__ ws, int64_t pld) {
    extern __shared__ __align__(16) unsigned char rsm2[];
    __shared__ ShearBatch B;
    {
        const int* src = reinterpret_cast<const int*>(batches + batch);
        int* dst = reinterpret_cast<int*>(&B);
        for (int i = threadIdx.x; i < (int)(sizeof(ShearBatch) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int njobs = B.njobs;
    uint4* stage_rec = reinterpret_cast<uint4*>(rsm2);                                                  // [SHR_ROWS][2]
    unsigned long long* stage_w = reinterpret_cast<unsigned long long*>(rsm2 + (size_t)SHR_ROWS * 32);  // [SHR_ROWS]
    unsigned* stage_pos = reinterpret_cast<unsigned*>(rsm2 + (size_t)SHR_ROWS * 40);                    // [SHR_ROWS]
    unsigned* lcur = reinterpret_cast<unsigned*>(rsm2 + (size_t)SHR_ROWS * 44);                         // [2][256]
    unsigned* gdel = lcur + 512;                                                                          // [njobs][256]
    unsigned* lstart = gdel + (size_t)njobs * 256;                                                        // [njobs][256]
    ShearLaneParam* lp = reinterpret_cast<ShearLaneParam*>(lstart + (size_t)njobs * 256);                 // [32]
    double* tile = reinterpret_cast<double*>(lp + 32);                                                    // [ncols][SHR_ROWS]
    unsigned char* b1s = reinterpret_cast<unsigned char*>(tile + (size_t)B.ncols * SHR_ROWS);             // [njobs][SHR_ROWS]
    const ShearRecJob* bj = jobs + B.job0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * SHR_ROWS;
    const int nrow = (int)min((int64_t)SHR_ROWS, N - r0);
    for (int i = threadIdx.x; i < njobs * 256; i += blockDim.x) gdel[i] = 0;
    for (int c = 0; c < B.ncols; c++) {
        const double* col = dX + (int64_t)B.cols[c] * ld + r0;
        for (int t = threadIdx.x; t < nrow; t += blockDim.x) tile[c * SHR_ROWS + t] = ldg_stream(col + t);
    }
    __syncthreads();
    const int t = threadIdx.x;  // SHR_ROWS == blockDim.x: one row per thread
    const bool has_row = t < nrow;
    unsigned skipmask = 0;
    // phase A: b1 of every anchor, rows per (anchor, bucket)
    if (has_row)
        for (int j = 0; j < njobs; j++) {
            const int b1 = bin_index_trunc_fx(tile[bj[j].acol * SHR_ROWS + t], bj[j].p1_min, bj[j].dx1, bj[j].inv1s);
            if (b1 < 0 || b1 > 255) skipmask |= 1u << j;
            const int c = min(max(b1, 0), 255);
            b1s[j * SHR_ROWS + t] = (unsigned char)c;
            atomicAdd(&gdel[j * 256 + c], 1u);
        }
    __syncthreads();
    // phase B: one warp per anchor: local starts + one global atomic per non-empty bucket (see k_bucket_records)
    for (int j = warp; j < njobs; j += nwarps) {
        unsigned v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[k] = gdel[j * 256 + lane * 8 + k];
            sum += v[k];
        }
        unsigned incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        unsigned pre = incl - sum;
        unsigned base[8];
#pragma unroll
        for (int k = 0; k < 8; k++) base[k] = v[k] ? atomicAdd(cursor + (size_t)j * 256 + lane * 8 + k, v[k]) : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            lstart[j * 256 + lane * 8 + k] = pre;
            gdel[j * 256 + lane * 8 + k] = base[k] - pre;
            pre += v[k];
        }
    }
    __syncthreads();
    const unsigned long long wrow_q = has_row ? wq[r0 + t] : 0ull;
    if (threadIdx.x < 256 && njobs > 0) lcur[threadIdx.x] = lstart[threadIdx.x];
    for (int j = 0; j < njobs; j++) {
        unsigned* cur = lcur + (j & 1) * 256;
        const ShearRecJob jb = bj[j];
        if (threadIdx.x < 32) {  // the partners' parameters of this anchor (read back after the barrier below)
            ShearLaneParam q{};
            if ((int)threadIdx.x < jb.np) {
                const ShearPairRef pr = pairs[B.pair0 + jb.pair0 + threadIdx.x];
                const ShearGeom g = geom[pr.job];
                q.r0 = pr.r0;
                q.r1 = pr.r1;
                q.rmin = g.rmin;
                q.dx = g.dx;
                q.invs = g.inv * 1048576.0;
                q.pcol = pr.pcol;
            }
            lp[threadIdx.x] = q;
        }
        __syncthreads();  // cursors + lane parameters of job j ready; copy-out of job j-1 done with the staging buffers
        if (has_row) {
            const unsigned c = b1s[j * SHR_ROWS + t];
            const unsigned li = atomicAdd(&cur[c], 1u);
            const double xi = tile[jb.acol * SHR_ROWS + t];
            unsigned v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                v[k] = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int l = 4 * k + b;
                    if (l < jb.np) {
                        const ShearLaneParam q = lp[l];
                        const double p2 = shear_p2(xi, tile[q.pcol * SHR_ROWS + t], q.r0, q.r1);
                        const int b2 = bin_index_trunc_fx(p2, q.rmin, q.dx, q.invs);
                        v[k] |= (unsigned)min(max(b2, 0), 255) << (8 * b);
                    }
                }
            }
            stage_rec[2 * li] = make_uint4(v[0], v[1], v[2], v[3]);
            stage_rec[2 * li + 1] = make_uint4(v[4], v[5], v[6], v[7]);
            stage_w[li] = ((skipmask >> j) & 1u) ? 0ull : wrow_q;
            stage_pos[li] = li + gdel[j * 256 + c];
        }
        __syncthreads();
        uint4* dstj = recs + (size_t)j * pld * 2;
        for (int i = threadIdx.x; i < 2 * nrow; i += blockDim.x) dstj[(size_t)stage_pos[i >> 1] * 2 + (i & 1)] = stage_rec[i];
        for (int i = threadIdx.x; i < nrow; i += blockDim.x) ws[(size_t)j * pld + stage_pos[i]] = stage_w[i];
        if (threadIdx.x < 256 && j + 1 < njobs) lcur[((j + 1) & 1) * 256 + threadIdx.x] = lstart[(j + 1) * 256 + threadIdx.x];
    }
}
