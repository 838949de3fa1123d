// kernels_stats.cuh -- N-sized kernels for the resident sample store and the weighted moments.
//
// Data layout in HBM: samples column-major, one contiguous, 512-byte aligned N-vector of
// float64 per parameter (dX[j*ld + n]); float64 weights dW[n]; 64-bit fixed-point weights dWq[n] =
// rint(w * 2^wshift) with sum(dWq) < 2^62 (exact, order-independent accumulation for histograms and
// order statistics).  Rows are cut into "segments" that never straddle a chain boundary.
#pragma once
#include <stdint.h>

struct Seg {
    int64_t r0, r1;
    int32_t chain;
    int32_t pad;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 16-byte load that does not allocate in L1 (each sample is touched once per sweep)
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldg_stream(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ ulonglong2 ldg_stream2_u64(const unsigned long long* p) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}

// ---- upload: row-major staging chunk [rows][P] -> column-major store --------------------------------
// 32x32 tile transpose through shared memory; reads coalesced along P, writes coalesced along N.
__global__ void k_transpose_in(const double* __restrict__ stage, int64_t rows, int P, double* __restrict__ dX,
                               int64_t ld, int64_t row0) {
    __shared__ double tile[32][33];
    const int64_t rb = (int64_t)blockIdx.x * 32;
    const int cb = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = rb + k;
        const int c = cb + tx;
        if (r < rows && c < P) tile[k][tx] = stage[r * P + c];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int c = cb + k;
        const int64_t r = rb + tx;
        if (r < rows && c < P) dX[(int64_t)c * ld + row0 + r] = tile[tx][k];
    }
}

// ---- weight statistics: per-block partials {sum w, sum w^2, max w, min w} ------------------------------
__global__ void k_wstats(const double* __restrict__ w, int64_t N, double* __restrict__ part) {
    double s = 0, s2 = 0, mx = -INFINITY, mn = INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = w[i];
        s += v;
        s2 += v * v;
        mx = fmax(mx, v);
        mn = fmin(mn, v);
    }
    __shared__ double sh[4][32];
    s = warp_sum(s);
    s2 = warp_sum(s2);
    mx = warp_max(mx);
    mn = warp_min(mn);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sh[0][wid] = s;
        sh[1][wid] = s2;
        sh[2][wid] = mx;
        sh[3][wid] = mn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        double a = 0, b = 0, c = -INFINITY, d = INFINITY;
        for (int i = 0; i < nw; i++) {
            a += sh[0][i];
            b += sh[1][i];
            c = fmax(c, sh[2][i]);
            d = fmin(d, sh[3][i]);
        }
        part[blockIdx.x * 4 + 0] = a;
        part[blockIdx.x * 4 + 1] = b;
        part[blockIdx.x * 4 + 2] = c;
        part[blockIdx.x * 4 + 3] = d;
    }
}

// fixed-point weights + count of outliers (w > mult_max, mcsamples.py:559-560) + exact integer total
__global__ void k_make_wq(const double* __restrict__ w, int64_t N, double scale, double mult_max,
                          unsigned long long* __restrict__ wq, unsigned long long* __restrict__ acc /*[2]*/) {
    unsigned long long tot = 0, outl = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = w[i];
        const unsigned long long q = (unsigned long long)__double2ll_rn(v * scale);
        wq[i] = q;
        tot += q;
        outl += (v > mult_max) ? 1ull : 0ull;
    }
    tot = warp_sum_u64(tot);
    outl = warp_sum_u64(outl);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[0], tot);  // integer: order independent
        atomicAdd(&acc[1], outl);
    }
}

// mean-likelihood weights (mcsamples.py:1556-1561, 1829-1831): pass 1 = per-block partials of sum(w * loglike)
// (chains.py:380-381), pass 2 = lw = w * exp(mean_loglike - loglike)
__global__ void k_wll_partial(const double* __restrict__ w, const double* __restrict__ ll, int64_t N, double* __restrict__ part) {
    double s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) s += w[i] * ll[i];
    __shared__ double sh[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) a += sh[i];
        part[blockIdx.x] = a;
    }
}
__global__ void k_like_weights(const double* __restrict__ w, const double* __restrict__ ll, double mean_loglike, int64_t N,
                               double* __restrict__ lw) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        lw[i] = w[i] * exp(mean_loglike - ll[i]);
}

__global__ void k_fill(double* p, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- per (segment, parameter): sum w x, sum w, min x, max x ----------------------------------------------
// grid (nseg, P); partial layout part[(seg*P + j)*4 + {0..3}]
__global__ void __launch_bounds__(256) k_col_sums(const double* __restrict__ dX, int64_t ld, const double* __restrict__ dW,
                                                  const Seg* __restrict__ segs, int P, double* __restrict__ part) {
    const Seg sg = segs[blockIdx.x];
    const int j = blockIdx.y;
    const double* x = dX + (int64_t)j * ld;
    double swx = 0, sw = 0, mn = INFINITY, mx = -INFINITY;
    // r0 is a multiple of 2 unless it is a chain boundary; handle an odd head element separately
    int64_t r = sg.r0;
    if ((r & 1) && r < sg.r1) {
        if (threadIdx.x == 0) {
            const double xv = x[r], wv = dW[r];
            swx += wv * xv;
            sw += wv;
            mn = fmin(mn, xv);
            mx = fmax(mx, xv);
        }
        r++;
    }
    const int64_t npair = (sg.r1 - r) >> 1;
    for (int64_t i = threadIdx.x; i < npair; i += blockDim.x) {
        const double2 xv = ldg_stream2(x + r + 2 * i);
        const double2 wv = *reinterpret_cast<const double2*>(dW + r + 2 * i);
        swx += wv.x * xv.x;
        swx += wv.y * xv.y;
        sw += wv.x + wv.y;
        mn = fmin(mn, fmin(xv.x, xv.y));
        mx = fmax(mx, fmax(xv.x, xv.y));
    }
    if (((sg.r1 - r) & 1) && threadIdx.x == 0) {
        const double xv = x[sg.r1 - 1], wv = dW[sg.r1 - 1];
        swx += wv * xv;
        sw += wv;
        mn = fmin(mn, xv);
        mx = fmax(mx, xv);
    }
    __shared__ double sh[4][8];
    swx = warp_sum(swx);
    sw = warp_sum(sw);
    mn = warp_min(mn);
    mx = warp_max(mx);
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sh[0][wid] = swx;
        sh[1][wid] = sw;
        sh[2][wid] = mn;
        sh[3][wid] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = INFINITY, d = -INFINITY;
        for (int i = 0; i < 8; i++) {
            a += sh[0][i];
            b += sh[1][i];
            c = fmin(c, sh[2][i]);
            d = fmax(d, sh[3][i]);
        }
        double* o = part + ((int64_t)blockIdx.x * P + j) * 4;
        o[0] = a;
        o[1] = b;
        o[2] = c;
        o[3] = d;
    }
}

// ---- centred second moments: S_c[i][j] = sum_{n in segment} w_n (x_ni - m_ci)(x_nj - m_cj) -----------------
// 64x64 parameter tile per CTA (upper-triangle tiles only), 256 threads, 4x4 outputs per thread.
// Rows are staged through shared memory 32 at a time, column-major with a padded leading dimension so
// that both the staging stores (lane = row) and the inner-product loads are bank-conflict free.
#define COV_T 64
#define COV_RB 32
__global__ void __launch_bounds__(256) k_cov_tiles(const double* __restrict__ dX, int64_t ld, const double* __restrict__ dW,
                                                   const Seg* __restrict__ segs, int P, int ntile,
                                                   const int2* __restrict__ tiles, const double* __restrict__ chain_means,
                                                   double* __restrict__ part /*[nseg][ntile][64*64]*/) {
    __shared__ double As[COV_T][COV_RB + 1];
    __shared__ double Bs[COV_T][COV_RB + 1];
    const Seg sg = segs[blockIdx.x];
    const int2 tl = tiles[blockIdx.y];
    const int i0 = tl.x * COV_T, j0 = tl.y * COV_T;
    const double* mean = chain_means + (int64_t)sg.chain * P;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0;
    for (int64_t rb = sg.r0; rb < sg.r1; rb += COV_RB) {
        const int64_t r = rb + lane;
        const bool ok = r < sg.r1;
        const double wv = ok ? dW[r] : 0.0;
        // 8 warps x 8 columns each per side
        for (int c = wid; c < COV_T; c += 8) {
            const int ci = i0 + c, cj = j0 + c;
            double a = 0, b = 0;
            if (ok && ci < P) a = (dX[(int64_t)ci * ld + r] - mean[ci]) * wv;
            if (ok && cj < P) b = dX[(int64_t)cj * ld + r] - mean[cj];
            As[c][lane] = a;
            Bs[c][lane] = b;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < COV_RB; k++) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; a++) av[a] = As[ty + 16 * a][k];
#pragma unroll
            for (int b = 0; b < 4; b++) bv[b] = Bs[tx + 16 * b][k];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
    double* o = part + ((int64_t)blockIdx.x * ntile + blockIdx.y) * (COV_T * COV_T);
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) o[(ty + 16 * a) * COV_T + tx + 16 * b] = acc[a][b];
}

// deterministic reduction of the per-segment tiles into per-chain matrices S[chain][P][P] (upper tiles,
// mirrored).  grid (ntile, nchains), 256 threads.
__global__ void k_cov_reduce(const double* __restrict__ part, const Seg* __restrict__ segs, int nseg, int ntile,
                             const int2* __restrict__ tiles, int P, double* __restrict__ S /*[nch][P][P]*/) {
    const int t = blockIdx.x, ch = blockIdx.y;
    const int2 tl = tiles[t];
    for (int e = threadIdx.x; e < COV_T * COV_T; e += blockDim.x) {
        double s = 0;
        for (int sgi = 0; sgi < nseg; sgi++)
            if (segs[sgi].chain == ch) s += part[((int64_t)sgi * ntile + t) * (COV_T * COV_T) + e];
        const int i = tl.x * COV_T + e / COV_T, j = tl.y * COV_T + e % COV_T;
        if (i < P && j < P) {
            double* Sc = S + (int64_t)ch * P * P;
            if (tl.x == tl.y) {
                if (j >= i) {
                    Sc[(int64_t)i * P + j] = s;
                    Sc[(int64_t)j * P + i] = s;
                }
            } else {
                Sc[(int64_t)i * P + j] = s;
                Sc[(int64_t)j * P + i] = s;
            }
        }
    }
}

// ---- lagged sums (MCMC effective-sample estimate) ----------------------------------------------------------
#define LAG_MAXK 16
struct LagJob {
    int param, mode;
    long long k0;
    int nk, pad;
    double mean, inv4s2;
};
// grid (nchunk, njobs), 256 threads; part[(job*nchunk + chunk)*LAG_MAXK + k].  Rows are NOT cut at chain
// boundaries: the reference neglects edge effects between concatenated chains (chains.py:427-428).
__global__ void __launch_bounds__(256) k_lag_sums(const double* __restrict__ dX, int64_t ld, const double* __restrict__ dW,
                                                  int64_t N, int64_t chunk, const LagJob* __restrict__ jobs,
                                                  double* __restrict__ part) {
    const LagJob jb = jobs[blockIdx.y];
    const double* x = dX + (int64_t)jb.param * ld;
    const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(N, r0 + chunk);
    double acc[LAG_MAXK];
#pragma unroll
    for (int k = 0; k < LAG_MAXK; k++) acc[k] = 0;
    for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
        const double xi = x[i], wi = dW[i];
        if (jb.mode == 0) {
            const double di = (xi - jb.mean) * wi;
#pragma unroll
            for (int k = 0; k < LAG_MAXK; k++) {
                const int64_t j = i + jb.k0 + k;
                if (k < jb.nk && j < N) acc[k] += di * ((x[j] - jb.mean) * dW[j]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < LAG_MAXK; k++) {
                const int64_t j = i + jb.k0 + k;
                if (k < jb.nk && j < N) {
                    const double d = xi - x[j];
                    acc[k] += (exp(-(d * d) * jb.inv4s2) * wi) * dW[j];
                }
            }
        }
    }
    __shared__ double sh[LAG_MAXK][8];
    const int wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < LAG_MAXK; k++) {
        const double v = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) sh[k][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < LAG_MAXK) {
        double t = 0;
        for (int i = 0; i < 8; i++) t += sh[threadIdx.x][i];
        part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * LAG_MAXK + threadIdx.x] = t;
    }
}
