// kernels_quant.cuh -- exact weighted order statistics by multi-pass radix selection.
//
// Replaces initParamConfidenceData + confidence (chains.py:793-838): the reference argsorts all N
// samples for every density; here all (parameter, fraction) targets are resolved together in a few
// streaming sweeps over the column-major store:
//   keys   : the IEEE-754 bits of x mapped to an order-preserving uint64
//   weights: 64-bit fixed point (exact integer sums -> the selected sample does not depend on the
//            summation order, unlike a float64 prefix sum)
//   pass   : every slot (param, target) keeps a key interval [klo, khi] known to contain the answer and
//            the cumulative weight below it; the keys inside the interval are histogrammed by their
//            leading bits (shared-memory privatised, native 32-bit integer atomics on two limbs),
//            scanned, and the interval is narrowed.  In the first pass all slots of a parameter share
//            the interval [key(min), key(max)] and therefore one histogram.
//   finish : an interval that collapses to one key is the answer; otherwise its candidates are gathered
//            (<= QCAP), bitonic-sorted in shared memory and scanned.  An interval with more candidates
//            than QCAP (heavy duplicates / clusters) is refined again.
#pragma once
#include <stdint.h>
#include <string.h>

#include "kernels_stats.cuh"

#define QMAXF 16   // max target fractions per parameter
#define QCAP 4096  // max candidates gathered per slot

struct QSlot {
    unsigned long long klo, khi;  // inclusive key interval
    unsigned long long below;     // fixed-point weight strictly below klo
    unsigned long long target;    // answer = first sample with inclusive cumulative weight >= target
    int shift;                    // digit = (key - klo) >> shift
    int state;                    // 0 = refining, 1 = resolved (value valid), 2 = ready to gather
    int ncand;                    // candidates gathered
    int pad;
    double value;
};

struct QCand {
    unsigned long long key, wq;
};

// first-pass geometry of a parameter: after pass 1 every slot interval lies inside ONE first-pass digit, so a
// table indexed by that digit (bit mask of slots) replaces the per-element scan over all slot intervals
struct QBase {
    unsigned long long klo0;
    int shift0, nb0;
};

__host__ __device__ __forceinline__ unsigned long long f64_to_key(double x) {
    unsigned long long b;
#if defined(__CUDA_ARCH__)
    b = (unsigned long long)__double_as_longlong(x);
#else
    memcpy(&b, &x, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double key_to_f64(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

__host__ __device__ __forceinline__ int q_shift_for(unsigned long long width, int log2bins) {
    int bits = 0;
    while (bits < 64 && (width >> bits)) bits++;  // bits needed to represent width
    const int sh = bits - log2bins;
    return sh > 0 ? sh : 0;
}

// 64-bit add into two 32-bit shared-memory limbs with native ATOMS.ADD (shared f64/u64 atomics compile to
// CAS spin loops on sm_100a; 32-bit integer adds are native)
__device__ __forceinline__ void smem_add_u64(unsigned* lo, unsigned* hi, unsigned long long v) {
    const unsigned vlo = (unsigned)v, vhi = (unsigned)(v >> 32);
    const unsigned old = atomicAdd(lo, vlo);
    const unsigned carry = (old + vlo < old) ? 1u : 0u;
    if (vhi | carry) atomicAdd(hi, vhi + carry);
}

// Same, addressed by 32-bit shared-window addresses (explicit atom.shared / red.shared: the generic-pointer form
// makes the compiler rebuild the shared window base around every update); the high limb lives `hi_off` bytes above
// the low limb.  add.cc/addc fold the carry of the low limb into the high-limb addend.  No "memory" clobber on purpose
// (it would make the compiler reload every shared-memory parameter after each update): the bins are only read back
// after a __syncthreads(), which is a compiler barrier.
__device__ __forceinline__ void smem_add_u64_addr(unsigned addr_lo, unsigned hi_off, unsigned long long v) {
    const unsigned vlo = (unsigned)v, vhi = (unsigned)(v >> 32);
    unsigned old, c;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr_lo), "r"(vlo));
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(c) : "r"(old), "r"(vlo), "r"(vhi));
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr_lo + hi_off), "r"(c));
}

// One histogram pass.  grid (nseg, nparams).
//   shared_first != 0: one histogram per parameter over slot 0's interval (first pass).
//   else             : each refining slot s owns bins [s*nbins, (s+1)*nbins).
// ghist[(p*QMAXF + s) * nbins + d] accumulates with u64 atomics (integer: order independent).
__global__ void __launch_bounds__(1024) k_qhist(const double* __restrict__ dX, int64_t ld,
                                               const unsigned long long* __restrict__ dWq, const Seg* __restrict__ segs,
                                               const int* __restrict__ params, const QSlot* __restrict__ slots, int ns,
                                               int nbins, int shared_first, unsigned long long* __restrict__ ghist,
                                               const QBase* __restrict__ qbase) {
    extern __shared__ unsigned qsm[];  // lo limbs [nh][nbins], hi limbs [nh][nbins], then the prefilter table
    __shared__ unsigned long long s_lo[QMAXF], s_w[QMAXF];
    __shared__ int s_shift[QMAXF], s_slot[QMAXF];
    __shared__ int n_act;
    const int p = blockIdx.y;
    const Seg sg = segs[blockIdx.x];
    const double* x = dX + (int64_t)params[p] * ld;
    const int nh = shared_first ? 1 : ns;
    unsigned* hlo = qsm;
    unsigned* hhi = qsm + nh * nbins;
    if (threadIdx.x == 0) {
        int na = 0;
        for (int s = 0; s < nh; s++) {
            const QSlot q = slots[p * QMAXF + s];
            bool act = q.state == 0;
            if (shared_first) {  // active if any slot of the parameter still refines
                act = false;
                for (int t = 0; t < ns; t++) act |= slots[p * QMAXF + t].state == 0;
            }
            if (act) {
                s_lo[na] = q.klo;
                s_w[na] = q.khi - q.klo;
                s_shift[na] = q.shift;
                s_slot[na] = s;
                na++;
            }
        }
        n_act = na;
    }
    const QBase qb = qbase[p];
    unsigned* tbl = qsm + 2 * nh * nbins;
    const int ntbl = shared_first ? 0 : qb.nb0;
    for (int i = threadIdx.x; i < 2 * nh * nbins + ntbl; i += blockDim.x) qsm[i] = 0;
    __syncthreads();
    const int na = n_act;
    if (na == 0) return;
    if (!shared_first) {
        if (threadIdx.x < na) atomicOr(&tbl[(int)((s_lo[threadIdx.x] - qb.klo0) >> qb.shift0)], 1u << threadIdx.x);
        __syncthreads();
    }
    unsigned hbase = (unsigned)__cvta_generic_to_shared(hlo);
    asm volatile("mov.u32 %0, %0;" : "+r"(hbase));  // opaque: the 32-bit shared address stays in a register
    const unsigned hoff = (unsigned)(nh * nbins) * 4u;
    for (int64_t r0 = sg.r0 + threadIdx.x; r0 < sg.r1; r0 += 4 * (int64_t)blockDim.x) {
        unsigned long long key[4], wq[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {  // unconditional loads on clamped rows: all eight are in flight before the first use
            const int64_t r = min(r0 + (int64_t)k * blockDim.x, sg.r1 - 1);
            key[k] = (unsigned long long)__double_as_longlong(__ldcs(x + r));
            wq[k] = shared_first ? __ldcs(dWq + r) : 0ull;  // first pass: every sample lands in the histogram
        }
#pragma unroll
        for (int k = 0; k < 4; k++) key[k] = (key[k] >> 63) ? ~key[k] : (key[k] | 0x8000000000000000ull);  // f64_to_key
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t r = r0 + (int64_t)k * blockDim.x;
            if (r >= sg.r1) continue;
            if (shared_first) {
                const unsigned long long d = key[k] - s_lo[0];
                if (d <= s_w[0]) {
                    const int bin = (int)(d >> s_shift[0]);
                    smem_add_u64_addr(hbase + ((unsigned)bin << 2), hoff, wq[k]);
                }
            } else {
                unsigned m = tbl[(int)((key[k] - qb.klo0) >> qb.shift0)];
                while (m) {
                    const int t = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned long long d = key[k] - s_lo[t];
                    if (d <= s_w[t]) {
                        const int bin = s_slot[t] * nbins + (int)(d >> s_shift[t]);
                        smem_add_u64_addr(hbase + ((unsigned)bin << 2), hoff, dWq[r]);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int t = 0; t < na; t++) {
        const int s = s_slot[t];
        for (int d = threadIdx.x; d < nbins; d += blockDim.x) {
            const unsigned long long v = ((unsigned long long)hhi[s * nbins + d] << 32) | hlo[s * nbins + d];
            if (v) atomicAdd(&ghist[((int64_t)(p * QMAXF + s)) * nbins + d], v);
        }
    }
}

// Scan: grid (nparams), one warp per slot.  Find the digit where the inclusive cumulative weight first
// reaches the target, narrow the interval.  next_state: 0 -> refine again with 2^next_log2bins bins,
// 2 -> gather candidates next.
__global__ void k_qscan(QSlot* __restrict__ slots, int ns, int nbins, int shared_first, int next_state,
                        int next_log2bins, const unsigned long long* __restrict__ ghist) {
    const int p = blockIdx.x;
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (s >= ns) return;
    QSlot q = slots[p * QMAXF + s];
    if (q.state != 0) return;
    const unsigned long long* h = ghist + ((int64_t)(p * QMAXF + (shared_first ? 0 : s))) * nbins;
    unsigned long long run = q.below;  // cumulative weight before the current chunk
    int found = -1;
    unsigned long long below_found = 0;
    int last_nz = -1;
    unsigned long long below_last = 0;
    for (int base = 0; base < nbins && found < 0; base += 32) {
        const int d = base + lane;
        const unsigned long long v = d < nbins ? h[d] : 0ull;
        unsigned long long inc = v;  // inclusive prefix within the chunk
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const bool hit = (v != 0) && (run + inc >= q.target);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        const unsigned nz = __ballot_sync(0xffffffffu, v != 0);
        if (nz) {
            const int l = 31 - __clz(nz);
            last_nz = base + l;
            below_last = run + __shfl_sync(0xffffffffu, inc - v, l);
        }
        if (m) {
            const int l = __ffs(m) - 1;
            found = base + l;
            below_found = run + __shfl_sync(0xffffffffu, inc - v, l);
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (found < 0) {  // target beyond the total (rounding): clamp to the last occupied digit
        found = last_nz;
        below_found = below_last;
    }
    if (lane == 0) {
        if (found < 0) {  // empty interval: cannot happen for consistent inputs; resolve to klo
            q.state = 1;
            q.value = key_to_f64(q.klo);
        } else {
            const unsigned long long lo = q.klo + ((unsigned long long)found << q.shift);
            unsigned long long hi = (q.shift >= 64) ? q.khi : lo + ((1ull << q.shift) - 1ull);
            if (hi > q.khi || hi < lo) hi = q.khi;
            q.klo = lo;
            q.khi = hi;
            q.below = below_found;
            if (lo == hi) {
                q.state = 1;
                q.value = key_to_f64(lo);
            } else {
                q.state = next_state;
                q.shift = q_shift_for(hi - lo, next_log2bins);
                q.ncand = 0;
            }
        }
        slots[p * QMAXF + s] = q;
    }
}

// Gather candidates of the slots in state 2.  grid (nseg, nparams).
__global__ void __launch_bounds__(256) k_qgather(const double* __restrict__ dX, int64_t ld,
                                                 const unsigned long long* __restrict__ dWq, const Seg* __restrict__ segs,
                                                 const int* __restrict__ params, QSlot* __restrict__ slots, int ns,
                                                 QCand* __restrict__ cand, const QBase* __restrict__ qbase) {
    extern __shared__ unsigned gtbl[];  // prefilter table, nb0 entries
    __shared__ unsigned long long s_lo[QMAXF], s_w[QMAXF];
    __shared__ int s_slot[QMAXF];
    __shared__ int n_act;
    const int p = blockIdx.y;
    const Seg sg = segs[blockIdx.x];
    const double* x = dX + (int64_t)params[p] * ld;
    if (threadIdx.x == 0) {
        int na = 0;
        for (int s = 0; s < ns; s++) {
            const QSlot q = slots[p * QMAXF + s];
            if (q.state == 2) {
                s_lo[na] = q.klo;
                s_w[na] = q.khi - q.klo;
                s_slot[na] = s;
                na++;
            }
        }
        n_act = na;
    }
    const QBase qb = qbase[p];
    for (int i = threadIdx.x; i < qb.nb0; i += blockDim.x) gtbl[i] = 0;
    __syncthreads();
    const int na = n_act;
    if (na == 0) return;
    if (threadIdx.x < na) atomicOr(&gtbl[(int)((s_lo[threadIdx.x] - qb.klo0) >> qb.shift0)], 1u << threadIdx.x);
    __syncthreads();
    for (int64_t r0 = sg.r0 + threadIdx.x; r0 < sg.r1; r0 += 4 * (int64_t)blockDim.x) {
        unsigned long long key[4];
#pragma unroll
        for (int k = 0; k < 4; k++)  // unconditional loads on clamped rows: all four in flight before the first use
            key[k] = (unsigned long long)__double_as_longlong(__ldcs(x + min(r0 + (int64_t)k * blockDim.x, sg.r1 - 1)));
#pragma unroll
        for (int k = 0; k < 4; k++) key[k] = (key[k] >> 63) ? ~key[k] : (key[k] | 0x8000000000000000ull);  // f64_to_key
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t r = r0 + (int64_t)k * blockDim.x;
            if (r >= sg.r1) continue;
            unsigned m = gtbl[(int)((key[k] - qb.klo0) >> qb.shift0)];
            while (m) {
                const int t = __ffs(m) - 1;
                m &= m - 1;
                if (key[k] - s_lo[t] <= s_w[t]) {
                    const int gs = p * QMAXF + s_slot[t];
                    const int idx = atomicAdd(&slots[gs].ncand, 1);
                    if (idx < QCAP) cand[(int64_t)gs * QCAP + idx] = QCand{key[k], dWq[r]};
                }
            }
        }
    }
}

// Select: grid (nparams * QMAXF), 512 threads, 64 KB dynamic shared memory.  Sort the gathered
// candidates by key, walk the inclusive cumulative weight to the target.  Overflowing slots go back to
// refinement (state 0) and are counted in *n_overflow.
__global__ void __launch_bounds__(512) k_qselect(QSlot* __restrict__ slots, int ns, const QCand* __restrict__ cand,
                                                 int log2bins, int* __restrict__ n_overflow) {
    extern __shared__ unsigned long long ssm[];  // keys[QCAP] then wq[QCAP]
    const int gs = blockIdx.x;
    if ((gs % QMAXF) >= ns) return;
    QSlot q = slots[gs];
    if (q.state != 2) return;
    const int n = q.ncand;
    if (n > QCAP) {
        if (threadIdx.x == 0) {
            q.state = 0;
            q.ncand = 0;
            q.shift = q_shift_for(q.khi - q.klo, log2bins);
            slots[gs] = q;
            atomicAdd(n_overflow, 1);
        }
        return;
    }
    int m = 1;
    while (m < n) m <<= 1;
    unsigned long long* keys = ssm;
    unsigned long long* wqs = ssm + QCAP;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        if (i < n) {
            const QCand c = cand[(int64_t)gs * QCAP + i];
            keys[i] = c.key;
            wqs[i] = c.wq;
        } else {
            keys[i] = ~0ull;
            wqs[i] = 0;
        }
    }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a = keys[i], b = keys[l];
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[l] = a;
                        const unsigned long long t = wqs[i];
                        wqs[i] = wqs[l];
                        wqs[l] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        unsigned long long cum = q.below;
        unsigned long long ans = n > 0 ? keys[n - 1] : q.klo;  // clamp to the last candidate
        for (int i = 0; i < n; i++) {
            cum += wqs[i];
            if (cum >= q.target) {
                ans = keys[i];
                break;
            }
        }
        q.state = 1;
        q.value = key_to_f64(ans);
        slots[gs] = q;
    }
}

// ---- helpers of the convergence tests (getConvergeTests, mcsamples.py:964-1034) ------------------------------------
// sum of n fixed-point weights (order statistics on a row range: confidence(..., start, end), chains.py:814-838)
__global__ void __launch_bounds__(256) k_sum_u64(const unsigned long long* __restrict__ w, int64_t n, unsigned long long* __restrict__ acc) {
    unsigned long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += w[i];
    s = warp_sum_u64(s);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(acc, s);
}

#define FRC_CHUNK 4096
// per-chunk totals of the fixed-point weights: grid ceil(N / FRC_CHUNK), 256 threads
__global__ void __launch_bounds__(256) k_chunk_sums_u64(const unsigned long long* __restrict__ w, int64_t N, unsigned long long* __restrict__ sums) {
    __shared__ unsigned long long sh[8];
    const int64_t r0 = (int64_t)blockIdx.x * FRC_CHUNK;
    unsigned long long s = 0;
    for (int t = threadIdx.x; t < FRC_CHUNK; t += blockDim.x)
        if (r0 + t < N) s += w[r0 + t];
    s = warp_sum_u64(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += sh[i];
        sums[blockIdx.x] = t;
    }
}

// one CTA per target: job = (chunk, target inside the chunk); first row of the chunk whose inclusive cumulative weight
// reaches the target (np.searchsorted(cumsum, target), side='left'); clamps to the last row of the chunk
__global__ void __launch_bounds__(256) k_fraction_rows(const unsigned long long* __restrict__ w, int64_t N,
                                                       const unsigned long long* __restrict__ jobs, long long* __restrict__ rows) {
    __shared__ unsigned long long pre[256];
    const int64_t r0 = (int64_t)jobs[2 * blockIdx.x] * FRC_CHUNK;
    const unsigned long long target = jobs[2 * blockIdx.x + 1];
    constexpr int PER = FRC_CHUNK / 256;
    unsigned long long v[PER], s = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int64_t r = r0 + (int64_t)threadIdx.x * PER + k;
        v[k] = r < N ? w[r] : 0ull;
        s += v[k];
    }
    pre[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {  // exclusive prefix over the 256 thread totals
        unsigned long long run = 0;
        for (int i = 0; i < 256; i++) {
            const unsigned long long t = pre[i];
            pre[i] = run;
            run += t;
        }
        rows[blockIdx.x] = min(r0 + FRC_CHUNK, N) - 1;  // default: target beyond the chunk (rounding) -> its last row
    }
    __syncthreads();
    unsigned long long run = pre[threadIdx.x];
    if (run < target || (target == 0 && threadIdx.x == 0)) {
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const unsigned long long nxt = run + v[k];
            if ((run < target || (target == 0 && k == 0 && threadIdx.x == 0)) && nxt >= target) {
                const int64_t r = r0 + (int64_t)threadIdx.x * PER + k;
                if (r < N) rows[blockIdx.x] = r;
            }
            run = nxt;
        }
    }
}
