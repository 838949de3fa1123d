// kernels_1d.cuh -- the 1D density path on the device.
//   k_hist1d : weighted fine-grid histograms of all requested parameters in one sweep of the column-major
//              sample store (_binSamples + np.bincount, mcsamples.py:1486-1498, 1554)
//   k_kde1d  : one CTA per density: TMA bulk load of the fine grid into shared memory, then
//              kde1d_core (bandwidth, kernel, convolution, corrections, normalisation)
#pragma once
#include "kde1d_core.cuh"
#include "kernels_quant.cuh"

// Exact restatement of ((x - binmin) / fine_width + 0.5).astype(int): the same three correctly rounded
// IEEE operations numpy performs, then truncation.  `inv` = 1/fine_width is only used to decide whether
// the cheap product (x - binmin) * inv can round differently from the true quotient: the two differ by at
// most a few ulp, so unless the value sits within 1e-9 of an integer the truncation is identical.
__device__ __forceinline__ int bin_index_round(double x, double binmin, double fine_width, double inv) {
    const double d = __dsub_rn(x, binmin);
    const double q = fma(d, inv, 0.5);  // approximate quotient + 0.5; only its distance to an integer matters
    const int i = __double2int_rd(q);
    const double fr = q - (double)i;
    if (fr > 1e-9 && fr < 1.0 - 1e-9) return (q < 0 && fr != 0) ? i + 1 : i;  // astype(int) truncates toward zero
    return (int)__dadd_rn(__ddiv_rn(d, fine_width), 0.5);
}
// kde.bin_samples (kde_bandwidth.py:85-87): truncating ((x - range_min) / dx).astype(int)
__device__ __forceinline__ int bin_index_trunc(double x, double rmin, double dx, double inv) {
    const double d = __dsub_rn(x, rmin);
    const double q = __dmul_rn(d, inv);
    const int i = __double2int_rd(q);
    const double fr = q - (double)i;
    if (fr > 1e-9 && fr < 1.0 - 1e-9) return (q < 0 && fr != 0) ? i + 1 : i;
    return (int)__ddiv_rn(d, dx);
}

// Same result as bin_index_trunc with the quotient evaluated in 32-bit fixed point (one DMUL + F2I): the low 20 bits
// of I are the fractional part, and only when it lies within 8 * 2^-20 of an integer -- or the conversion saturated
// (d < 0 -> 0, huge -> 0xffffffff) -- is the exact IEEE division executed, out of line.  The approximation error of
// d * inv is < 2^-40 bins for indices below 4096, so the fast path never disagrees with the division.
__device__ __noinline__ int bin_index_trunc_exact(double d, double dx) { return (int)__ddiv_rn(d, dx); }
__device__ __forceinline__ int bin_index_trunc_fx(double x, double rmin, double dx, double inv_scaled /* 2^20 / dx */) {
    const double d = __dsub_rn(x, rmin);
    const unsigned I = __double2uint_rd(__dmul_rn(d, inv_scaled));
    if (__builtin_expect(((I & 0xfffffu) - 8u) >= (0xfffffu - 15u), 0)) return bin_index_trunc_exact(d, dx);
    return (int)(I >> 20);
}

// ---- TMA (bulk async copy) helpers: global -> shared with an mbarrier ------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(phase)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

struct Hist1dJob {
    int param, F;
    double binmin, fine_width, inv_width;
};

// grid (nseg, njobs), 256 threads, dynamic smem = 2 * F * 4 bytes (two 32-bit limbs per bin).
// Privatised shared-memory bins with native 32-bit integer atomics on 64-bit fixed-point weights;
// flushed with u64 global atomics (integer: the result is independent of the order of accumulation).
__global__ void __launch_bounds__(256) k_hist1d(const double* __restrict__ dX, int64_t ld,
                                                const unsigned long long* __restrict__ dWq, const Seg* __restrict__ segs,
                                                const Hist1dJob* __restrict__ jobs, unsigned long long* __restrict__ gbins,
                                                int64_t gstride) {
    extern __shared__ unsigned hsm[];
    const Hist1dJob jb = jobs[blockIdx.y];
    const Seg sg = segs[blockIdx.x];
    const int F = jb.F;
    unsigned* hlo = hsm;
    unsigned* hhi = hsm + F;
    for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) hsm[i] = 0;
    __syncthreads();
    const double* x = dX + (int64_t)jb.param * ld;
    int64_t r = sg.r0;
    if ((r & 1) && r < sg.r1) {
        if (threadIdx.x == 0) {
            const int b = bin_index_round(x[r], jb.binmin, jb.fine_width, jb.inv_width);
            if (b >= 0 && b < F) smem_add_u64(hlo + b, hhi + b, dWq[r]);
        }
        r++;
    }
    const int64_t npair = (sg.r1 - r) >> 1;
    // four independent 16-byte loads of x and of w in flight per thread before any dependent work
    for (int64_t i0 = threadIdx.x; i0 < npair; i0 += 4 * (int64_t)blockDim.x) {
        double2 xv[4];
        ulonglong2 wv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t i = i0 + (int64_t)k * blockDim.x;
            if (i < npair) {
                xv[k] = ldg_stream2(x + r + 2 * i);
                wv[k] = ldg_stream2_u64(dWq + r + 2 * i);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t i = i0 + (int64_t)k * blockDim.x;
            if (i < npair) {
                const int b0 = bin_index_round(xv[k].x, jb.binmin, jb.fine_width, jb.inv_width);
                const int b1 = bin_index_round(xv[k].y, jb.binmin, jb.fine_width, jb.inv_width);
                if (b0 >= 0 && b0 < F) smem_add_u64(hlo + b0, hhi + b0, wv[k].x);
                if (b1 >= 0 && b1 < F) smem_add_u64(hlo + b1, hhi + b1, wv[k].y);
            }
        }
    }
    if (((sg.r1 - r) & 1) && threadIdx.x == 0) {
        const int b = bin_index_round(x[sg.r1 - 1], jb.binmin, jb.fine_width, jb.inv_width);
        if (b >= 0 && b < F) smem_add_u64(hlo + b, hhi + b, dWq[sg.r1 - 1]);
    }
    __syncthreads();
    unsigned long long* g = gbins + (int64_t)blockIdx.y * gstride;
    for (int i = threadIdx.x; i < F; i += blockDim.x) {
        const unsigned long long v = ((unsigned long long)hhi[i] << 32) | hlo[i];
        if (v) atomicAdd(g + i, v);
    }
}

// ---- TMA-pipelined variant of the 1D sweep ------------------------------------------------------------------
// The sample and weight streams are staged through a ring of H1_STAGES shared-memory buffers by bulk async
// copies (cp.async.bulk -> UBLKCP) with full/empty mbarriers: no load latency on the consumers' critical path
// and no registers tied up by loads in flight.  Consumers copy their elements to registers, release the stage,
// then do the index arithmetic and the shared-memory atomics.  Bin index: q = (x - binmin)/fw + 0.5 is evaluated
// in fixed point (one DFMA + F2I); the low `sh` bits are the fractional part, and only when it lies within
// 2^-(sh-4) of an integer is the exact IEEE sequence (sub, div, add) executed.
#define H1_STAGES 3
#define H1_CHUNK 1024
// exact restatement of int((x - binmin)/fine_width + 0.5) given d = x - binmin (kept out of line: executed for
// about one sample in 30 000)
__device__ __noinline__ unsigned bin_index_exact(double d, double fine_width) {
    return (unsigned)(int)__dadd_rn(__ddiv_rn(d, fine_width), 0.5);
}

struct Hist1dJobT {
    int param, F, sh, pad;
    double binmin, fine_width, inv_width, scale;  // scale = 2^sh
};

__global__ void __launch_bounds__(256) k_hist1d_tma(const double* __restrict__ dX, int64_t ld,
                                                    const unsigned long long* __restrict__ dWq, const Seg* __restrict__ segs,
                                                    const Hist1dJobT* __restrict__ jobs, unsigned long long* __restrict__ gbins,
                                                    int64_t gstride) {
    extern __shared__ __align__(128) unsigned char tsm[];
    __shared__ __align__(8) unsigned long long full[H1_STAGES], empty[H1_STAGES];
    const Hist1dJobT jb = jobs[blockIdx.y];
    const Seg sg = segs[blockIdx.x];
    const int F = jb.F;
    double* xs = reinterpret_cast<double*>(tsm);                                                   // [stages][chunk]
    unsigned long long* ws = reinterpret_cast<unsigned long long*>(tsm + (size_t)H1_STAGES * H1_CHUNK * 8);
    unsigned* hlo = reinterpret_cast<unsigned*>(tsm + (size_t)2 * H1_STAGES * H1_CHUNK * 8);
    unsigned* hhi = hlo + F;
    for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) hlo[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < H1_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], blockDim.x);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const double* x = dX + (int64_t)jb.param * ld + sg.r0;  // r0 even -> 16-byte aligned
    const unsigned long long* wq = dWq + sg.r0;
    const int64_t n = sg.r1 - sg.r0;
    const int nchunks = (int)((n + H1_CHUNK - 1) / H1_CHUNK);
    auto issue = [&](int c) {
        const int s = c % H1_STAGES;
        const int64_t off = (int64_t)c * H1_CHUNK;
        const int cnt = (int)min((int64_t)H1_CHUNK, n - off);
        const unsigned bytes = (unsigned)(((cnt + 1) & ~1) * 8);  // columns are padded: one element past is readable
        mbar_expect_tx(&full[s], 2 * bytes);
        bulk_g2s(xs + (size_t)s * H1_CHUNK, x + off, bytes, &full[s]);
        bulk_g2s(ws + (size_t)s * H1_CHUNK, wq + off, bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (int c = 0; c < H1_STAGES && c < nchunks; c++) issue(c);
    const unsigned fmask = (1u << jb.sh) - 1u;
    const double kscale = jb.inv_width * jb.scale, khalf = 0.5 * jb.scale;
    unsigned hbase = (unsigned)__cvta_generic_to_shared(hlo);
    asm volatile("mov.u32 %0, %0;" : "+r"(hbase));  // opaque: the 32-bit shared address stays in a register
    const unsigned hoff = (unsigned)F * 4u;
    for (int c = 0; c < nchunks; c++) {
        const int s = c % H1_STAGES;
        const unsigned ph = (unsigned)((c / H1_STAGES) & 1);
        mbar_wait(&full[s], ph);
        const int cnt = (int)min((int64_t)H1_CHUNK, n - (int64_t)c * H1_CHUNK);
        const double2* xp = reinterpret_cast<const double2*>(xs + (size_t)s * H1_CHUNK);
        const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(ws + (size_t)s * H1_CHUNK);
        double2 xv[2];
        ulonglong2 wv[2];
#pragma unroll
        for (int k = 0; k < 2; k++) {  // 256 threads x 2 x (2 elements) = one chunk
            const int i = threadIdx.x + k * 256;
            xv[k] = xp[i];
            wv[k] = wp[i];
        }
        mbar_arrive(&empty[s]);  // stage may be refilled
        if (threadIdx.x == 0 && c + H1_STAGES < nchunks) {
            mbar_wait(&empty[s], ph);
            issue(c + H1_STAGES);
        }
        const bool fullc = cnt == H1_CHUNK;  // uniform: only the last chunk of a segment can be partial
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int i2 = (threadIdx.x + k * 256) * 2;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if (!fullc && i2 + e >= cnt) continue;
                const double xe = e ? xv[k].y : xv[k].x;
                const unsigned long long we = e ? wv[k].y : wv[k].x;
                const double d = __dsub_rn(xe, jb.binmin);
                // the conversion saturates: t < 0 -> 0 (fraction 0) and t >= 2^32 -> 0xffffffff (fraction all ones);
                // both fail the fraction test and take the exact path, so no separate range test is needed
                const unsigned I = __double2uint_rd(fma(d, kscale, khalf));
                unsigned b = I >> jb.sh;
                if (__builtin_expect(((I & fmask) - 8u) >= (fmask - 15u), 0)) b = bin_index_exact(d, jb.fine_width);
                if (b < (unsigned)F) smem_add_u64_addr(hbase + (b << 2), hoff, we);
            }
        }
    }
    __syncthreads();
    unsigned long long* g = gbins + (int64_t)blockIdx.y * gstride;
    for (int i = threadIdx.x; i < F; i += blockDim.x) {
        const unsigned long long v = ((unsigned long long)hhi[i] << 32) | hlo[i];
        if (v) atomicAdd(g + i, v);
    }
}

// fixed point -> float64 bins
__global__ void k_bins_to_f64(const unsigned long long* __restrict__ g, double* __restrict__ out, int64_t n, double inv_scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)g[i] * inv_scale;
}

struct Kde1dTables {
    const cplx* tw;
    const cplx* tw4;
    const double* cos4;
};

// grid (n), 256 threads.  Work arrays: 9*F doubles (11*F with mean likelihoods), in dynamic shared memory when use_smem != 0, otherwise in
// the per-density global workspace gwork + i*9*F.
__global__ void __launch_bounds__(256) k_kde1d(const gdk_spec1d* __restrict__ specs, const unsigned long long* __restrict__ gbins,
                                               int64_t gstride, double inv_scale, IsjConsts K, const Kde1dTables* __restrict__ tabs,
                                               double* __restrict__ P_out, int64_t pstride, gdk_result1d* __restrict__ res,
                                               double* __restrict__ gwork, int use_smem,
                                               const unsigned long long* __restrict__ gbins_l, double inv_scale_l,
                                               double* __restrict__ L_out) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ double red[32];
    __shared__ __align__(8) unsigned long long bar;
    const int i = blockIdx.x;
    const gdk_spec1d sp = specs[i];
    const int F = sp.fine_bins;
    const int nwork = gbins_l ? 11 : 9;
    double* base = use_smem ? reinterpret_cast<double*>(dsm) : gwork + (int64_t)i * nwork * F;
    Kde1dWork W;
    W.bins = base;
    W.a2 = base + F;
    W.logI = base + 2 * F;
    W.aux = base + 3 * F;
    W.aux2 = base + 4 * F;
    W.ca = reinterpret_cast<cplx*>(base + 5 * F);
    W.cb = reinterpret_cast<cplx*>(base + 7 * F);
    W.P = base + 5 * F;    // aliases ca (free after the DCT)
    W.win = base + 7 * F;  // aliases cb
    W.tw = tabs[i].tw;
    W.tw4 = tabs[i].tw4;
    W.cos4 = tabs[i].cos4;
    const unsigned long long* g = gbins + (int64_t)i * gstride;
    if (use_smem && ((F * 8) % 16 == 0)) {
        // TMA staging of the fine grid: one bulk copy of the raw fixed-point bins, completion on an mbarrier
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, (unsigned)(F * 8));
            bulk_g2s(W.aux, g, (unsigned)(F * 8), &bar);
        }
        mbar_wait(&bar, 0);
        const unsigned long long* raw = reinterpret_cast<const unsigned long long*>(W.aux);
        for (int k = threadIdx.x; k < F; k += blockDim.x) W.bins[k] = (double)raw[k] * inv_scale;
    } else {
        for (int k = threadIdx.x; k < F; k += blockDim.x) W.bins[k] = (double)g[k] * inv_scale;
    }
    if (gbins_l) {  // mean-likelihood histogram (meanlikes=True)
        double* lb = base + 9 * F;
        const unsigned long long* gl = gbins_l + (int64_t)i * gstride;
        for (int k = threadIdx.x; k < F; k += blockDim.x) lb[k] = (double)gl[k] * inv_scale_l;
        W.likebins = lb;
        W.raw = base + 10 * F;
        W.likes_out = L_out + (int64_t)i * pstride;
    }
    __syncthreads();
    CoopBlock co{(int)threadIdx.x, (int)blockDim.x, red};
    kde1d_core(co, sp, K, W, P_out + (int64_t)i * pstride, res + i);
}

// ---- raw ND histogram (getRawNDDensityGridData, mcsamples.py:2098-2166: _binSamples per axis + _makeNDhist) ----------
#define HND_MAXD 8
struct HistNdJob {
    int ndim, pad;
    int param[HND_MAXD], n[HND_MAXD];
    long long stride[HND_MAXD];  // axis 0 is the fastest index (flat = sum ix_d * stride_d, mcsamples.py:2034-2046)
    double binmin[HND_MAXD], fw[HND_MAXD], inv[HND_MAXD];
};
// grid (nseg), 512 threads.  use_smem: privatised two-limb bins in shared memory (total * 8 bytes), flushed with u64
// reductions; otherwise straight L2 reductions.  ll != NULL: profile likelihood -- per-bin maximum of
// exp(shift - loglike) kept as the bit pattern of a non-negative double (order preserving) with atomicMax.
__global__ void __launch_bounds__(512) k_histnd(const double* __restrict__ dX, int64_t ld, const unsigned long long* __restrict__ wq,
                                                const double* __restrict__ ll, double shift, const Seg* __restrict__ segs, HistNdJob jb,
                                                int total, int use_smem, unsigned long long* __restrict__ gbins) {
    extern __shared__ unsigned nsm[];
    if (use_smem) {
        for (int i = threadIdx.x; i < 2 * total; i += blockDim.x) nsm[i] = 0;
        __syncthreads();
    }
    const Seg sg = segs[blockIdx.x];
    for (int64_t r = sg.r0 + threadIdx.x; r < sg.r1; r += blockDim.x) {
        long long flat = 0;
        bool ok = true;
        for (int d = 0; d < jb.ndim; d++) {
            const int b = bin_index_round(ldg_stream(dX + (int64_t)jb.param[d] * ld + r), jb.binmin[d], jb.fw[d], jb.inv[d]);
            ok = ok && b >= 0 && b < jb.n[d];
            flat += (long long)b * jb.stride[d];
        }
        if (!ok) continue;
        if (ll) {
            const double v = exp(shift - ll[r]);
            atomicMax(gbins + flat, (unsigned long long)__double_as_longlong(v));
        } else if (use_smem) {
            smem_add_u64(nsm + flat, nsm + total + flat, wq[r]);
        } else {
            const unsigned long long w = wq[r];
            if (w) atomicAdd(gbins + flat, w);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const unsigned long long v = ((unsigned long long)nsm[total + i] << 32) | nsm[i];
            if (v) atomicAdd(gbins + i, v);
        }
    }
}
