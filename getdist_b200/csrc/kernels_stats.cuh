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

// ---- fused one-sweep weighted statistics (chains.py:373-412, 709-733, 1446-1474; mcsamples.py:552-576) -------------
// ONE pass over the samples gives, per row segment: sum w, sum w d, min x, max x and the P x P block sum w d d^T with
// d = x - s, s = the segment's own first row (a shift that costs nothing to obtain, is the same whatever the number of
// ranks or the upload chunking, and keeps the one-pass second moments free of cancellation: |d| is a few sigma).
// Segments are cut at absolute multiples of ST_SEG rows and at chain boundaries; they are merged in a fixed order --
// sequentially inside "stat blocks" (ST_BLOCK rows cut at chain boundaries) on the device, then sequentially over the
// blocks on the host -- with the pairwise update of the mean and the centred second moments
//     D = m_b - m_a,  W = A_a + A_b,  m = m_a + D A_b / W,  S = S_a + S_b + D D^T A_a A_b / W,
// so the result does not depend on which rank or which upload chunk a block was computed in.
//
// k_stats_fused: grid (segments, upper-triangle 64 x 64 parameter tiles), 128 threads, 4 x 8 outputs per thread.
// Rows are staged 32 at a time through shared memory (k-major, XOR-swizzled pairs: conflict-free 64-bit staging
// stores and conflict-free LDS.128 operand loads, 6 LDS.128 per 32 DFMA).  Diagonal tiles also produce the column sums
// and min/max (a row with weight 0 still counts for min/max, as in the reference), tile 0 the weight sum.
#define ST_T 64
#define ST_RB 32
#define ST_SEG 2048
#define ST_BLOCK 65536
#define ST_SMEM ((3 * ST_RB * ST_T + 2 * ST_T) * 8)
__device__ __forceinline__ int st_swz(int c, int k) { return ((((c >> 1) ^ (k & 15)) << 1) | (c & 1)); }

// per-segment partial: [ntile][64*64] second-moment tiles, then B[T*64], min[T*64], max[T*64], then A (+pad to even)
__host__ __device__ __forceinline__ int64_t st_part_stride(int T, int ntile) { return (int64_t)ntile * (ST_T * ST_T) + 3 * T * ST_T + 2; }

__global__ void __launch_bounds__(128, 3) k_stats_fused(const double* __restrict__ dX, int64_t ld, const double* __restrict__ dW,
                                                        const Seg* __restrict__ segs, int P, int T, int ntile,
                                                        const int2* __restrict__ tiles, double* __restrict__ part) {
    extern __shared__ __align__(16) double st_sm[];  // ST_SMEM bytes (opt-in: more than 48 KB)
    double(*As)[ST_T] = reinterpret_cast<double(*)[ST_T]>(st_sm);
    double(*Bs)[ST_T] = reinterpret_cast<double(*)[ST_T]>(st_sm + ST_RB * ST_T);
    double(*Xs)[ST_T] = reinterpret_cast<double(*)[ST_T]>(st_sm + 2 * ST_RB * ST_T);  // diagonal tiles: the raw values
                                                            // (exact min / max: (x - s) + s need not give x back)
    double* shA = st_sm + 3 * ST_RB * ST_T;
    double* shB = shA + ST_T;
    const Seg sg = segs[blockIdx.x];
    const int2 tl = tiles[blockIdx.y];
    const int i0 = tl.x * ST_T, j0 = tl.y * ST_T;
    const bool diag = tl.x == tl.y;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int tx = t & 7, ty = t >> 3;  // 8 x 16 threads, 4 (rows) x 8 (columns) outputs each
    if (t < ST_T) {
        const int ci = i0 + t, cj = j0 + t;
        shA[t] = ci < P ? dX[(int64_t)ci * ld + sg.r0] : 0.0;
        shB[t] = cj < P ? dX[(int64_t)cj * ld + sg.r0] : 0.0;
    }
    __syncthreads();
    double acc[4][8];
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
        for (int f = 0; f < 8; f++) acc[e][f] = 0;
    const int ct = t & (ST_T - 1);  // threads 0..63: column ct of a diagonal tile
    double sa = 0, mn = shA[ct], mx = shA[ct], sw = 0;
    for (int64_t rb = sg.r0; rb < sg.r1; rb += ST_RB) {
        const int64_t r = rb + lane;
        const bool ok = r < sg.r1;
        const int64_t rr = ok ? r : sg.r1 - 1;
        const double wv = ok ? dW[r] : 0.0;
        sw += wv;
        // warp `wid` stages columns [16 wid, 16 wid + 16) of both sides; lane = row (coalesced along N).  Loads are
        // unconditional on clamped addresses and issued eight at a time before their first use (the latency of a batch
        // is paid once); the out-of-range ones are replaced afterwards.
        if (diag) {
#pragma unroll
            for (int c8 = 0; c8 < 16; c8 += 8) {
                double xv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) xv[u] = __ldcs(dX + (int64_t)min(i0 + 16 * wid + c8 + u, P - 1) * ld + rr);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int c = 16 * wid + c8 + u;
                    const double xa = (ok && i0 + c < P) ? xv[u] : shA[c];
                    const double da = xa - shA[c];
                    const int sc = st_swz(c, lane);
                    As[lane][sc] = da * wv;
                    Bs[lane][sc] = da;
                    Xs[lane][sc] = xa;
                }
            }
        } else {
#pragma unroll
            for (int c4 = 0; c4 < 16; c4 += 4) {
                double xv[4], yv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    xv[u] = __ldcs(dX + (int64_t)min(i0 + 16 * wid + c4 + u, P - 1) * ld + rr);
                    yv[u] = __ldcs(dX + (int64_t)min(j0 + 16 * wid + c4 + u, P - 1) * ld + rr);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int c = 16 * wid + c4 + u;
                    const double da = (ok && i0 + c < P) ? xv[u] - shA[c] : 0.0;
                    const double db = (ok && j0 + c < P) ? yv[u] - shB[c] : 0.0;
                    const int sc = st_swz(c, lane);
                    As[lane][sc] = da * wv;
                    Bs[lane][sc] = db;
                }
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < ST_RB; k++) {
            double a[4], b[8];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const double2 av = *reinterpret_cast<const double2*>(&As[k][((ty + 16 * q) ^ (k & 15)) << 1]);
                a[2 * q] = av.x;
                a[2 * q + 1] = av.y;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double2 bv = *reinterpret_cast<const double2*>(&Bs[k][((tx + 8 * q) ^ (k & 15)) << 1]);
                b[2 * q] = bv.x;
                b[2 * q + 1] = bv.y;
            }
#pragma unroll
            for (int e = 0; e < 4; e++)
#pragma unroll
                for (int f = 0; f < 8; f++) acc[e][f] = fma(a[e], b[f], acc[e][f]);
        }
        if (diag && t < ST_T) {  // column t: sum of w d, exact min / max of x over the 32 staged rows
#pragma unroll 8
            for (int k = 0; k < ST_RB; k++) {
                sa += As[k][st_swz(t, k)];
                const double xv = Xs[k][st_swz(t, k)];
                mn = fmin(mn, xv);
                mx = fmax(mx, xv);
            }
        }
        __syncthreads();
    }
    double* o = part + (int64_t)blockIdx.x * st_part_stride(T, ntile);
    double* oc = o + (int64_t)blockIdx.y * (ST_T * ST_T);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int i = 2 * ty + (e & 1) + 32 * (e >> 1);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = 2 * tx + 16 * q;
            *reinterpret_cast<double2*>(&oc[i * ST_T + j]) = make_double2(acc[e][2 * q], acc[e][2 * q + 1]);
        }
    }
    if (diag) {
        double* ob = o + (int64_t)ntile * (ST_T * ST_T);
        if (t < ST_T) {
            ob[i0 + t] = sa;
            ob[T * ST_T + i0 + t] = mn;
            ob[2 * T * ST_T + i0 + t] = mx;
        }
        if (blockIdx.y == 0 && wid == 0) {
            sw = warp_sum(sw);
            if (lane == 0) ob[3 * T * ST_T] = sw;
        }
    }
}

// Sequential merge of the segments of one stat block: grid (blocks, slices), 256 threads, dynamic shared 3 P doubles.
// Every slice repeats the (cheap) recursion of the running mean and applies it to its own share of the P x P elements.
// Block record: [0] = A, [1 .. P] mean, [P+1 .. 2P] min x, [2P+1 .. 3P] max x, [3P+1 ..] S (P x P, symmetric, centred).
__host__ __device__ __forceinline__ int64_t st_block_stride(int P) { return 3 * (int64_t)P + 1 + (int64_t)P * P; }
__global__ void __launch_bounds__(256) k_stats_merge(const double* __restrict__ part, const Seg* __restrict__ segs,
                                                     const int2* __restrict__ blkseg /* first seg, count */,
                                                     const int* __restrict__ blkout, int seg0, const double* __restrict__ dX, int64_t ld,
                                                     int P, int T, int ntile, double* __restrict__ bout) {
    extern __shared__ double shm[];
    double* m = shm;           // running mean
    double* dl = shm + P;      // D = m_seg - m
    double* ds = shm + 2 * P;  // delta = B_seg / A_seg
    const int2 bs = blkseg[blockIdx.x];
    double* out = bout + (int64_t)blkout[blockIdx.x] * st_block_stride(P);
    double* S = out + 3 * P + 1;
    const int64_t pstride = st_part_stride(T, ntile);
    const bool first = blockIdx.y == 0;  // slice 0 also writes the vectors
    for (int c = threadIdx.x; c < P; c += blockDim.x) {
        m[c] = 0;
        if (first) {
            out[1 + P + c] = INFINITY;
            out[1 + 2 * P + c] = -INFINITY;
        }
    }
    const int64_t e0 = (int64_t)blockIdx.y * blockDim.x + threadIdx.x, estep = (int64_t)gridDim.y * blockDim.x;
    for (int64_t e = e0; e < (int64_t)P * P; e += estep) S[e] = 0;
    double A = 0;
    __syncthreads();
    for (int s = bs.x; s < bs.x + bs.y; s++) {
        const double* ps = part + (int64_t)(s - seg0) * pstride;  // the partial buffer starts at segment seg0
        const double* pb = ps + (int64_t)ntile * (ST_T * ST_T);
        const double As_ = pb[3 * T * ST_T];
        const int64_t r0 = segs[s].r0;
        for (int c = threadIdx.x; c < P; c += blockDim.x) {
            const double sh = dX[(int64_t)c * ld + r0];
            if (first) {
                out[1 + P + c] = fmin(out[1 + P + c], pb[T * ST_T + c]);
                out[1 + 2 * P + c] = fmax(out[1 + 2 * P + c], pb[2 * T * ST_T + c]);
            }
            if (As_ > 0) {
                const double d = pb[c] / As_;
                ds[c] = d;
                dl[c] = (A > 0) ? (sh + d) - m[c] : 0.0;
                if (!(A > 0)) m[c] = sh + d;
            }
        }
        __syncthreads();
        if (As_ > 0) {
            const double W = A + As_;
            const double f = (A > 0) ? A * As_ / W : 0.0;
            for (int64_t e = e0; e < (int64_t)P * P; e += estep) {
                const int i = (int)(e / P), j = (int)(e % P);
                const int a = min(i, j), b = max(i, j);
                const int ta = a / ST_T, tb = b / ST_T;
                const int tix = ta * T - (ta * (ta - 1)) / 2 + (tb - ta);
                const double cv = ps[(int64_t)tix * (ST_T * ST_T) + (a % ST_T) * ST_T + (b % ST_T)];
                S[e] += (cv - As_ * ds[i] * ds[j]) + f * dl[i] * dl[j];
            }
            __syncthreads();
            if (A > 0)
                for (int c = threadIdx.x; c < P; c += blockDim.x) m[c] += dl[c] * (As_ / W);
            A = W;
        }
        __syncthreads();
    }
    if (first) {
        for (int c = threadIdx.x; c < P; c += blockDim.x) out[1 + c] = m[c];
        if (threadIdx.x == 0) out[0] = A;
    }
}

// ---- multi-GPU: finished result grids of this rank stored into every peer's gathered window over NVLink ------------
// (peer memory mapped with CUDA IPC; launched on the second stream behind the group-finished event, so the transfer of
// one group overlaps the convolutions of the next).  grid (densities, peers); 16-byte stores when aligned.
struct PeerTable {
    double* base[16];
    int n;
};
__global__ void __launch_bounds__(256) k_push_peers(const double* __restrict__ local, const long long* __restrict__ offs,
                                                    const int* __restrict__ counts, long long fixed_count, PeerTable peers) {
    const long long off = offs[blockIdx.x];
    const long long cnt = counts ? counts[blockIdx.x] : fixed_count;
    const double* src = local + off;
    double* dst = peers.base[blockIdx.y] + off;
    if (((off | cnt) & 1) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(src);
        double2* d2 = reinterpret_cast<double2*>(dst);
        for (long long i = threadIdx.x; i < cnt / 2; i += blockDim.x) d2[i] = s2[i];
    } else {
        for (long long i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = src[i];
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
