// coop.cuh -- "cooperative group of threads" abstraction used by all grid-stage device code.
//
// Every grid-sized routine (FFT, DCT, root finders, convolution, corrections) is written once as a
// template over a Coop type:
//   co.tid, co.nt        this thread's index / number of cooperating threads
//   co.sync()            barrier over the group
//   co.sum(v), co.max(v) reduction over the group, result broadcast to every thread (deterministic order)
// On the device Coop = CoopBlock (one CTA: warp shuffles + one shared-memory hop).  On the host
// Coop = CoopHost (tid 0 of 1, barriers are no-ops) which lets tests/hostsim compile the very same
// headers with g++ and check the arithmetic against the oracle in a container without a GPU.
// The host instantiation is TEST INFRASTRUCTURE; the product never runs it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GDK_HD __host__ __device__ __forceinline__
#define GDK_D __device__ __forceinline__
#else
#define GDK_HD inline
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct cplx {
    double x, y;
};
GDK_HD cplx cmul(cplx a, cplx b) { return cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
GDK_HD cplx cadd(cplx a, cplx b) { return cplx{a.x + b.x, a.y + b.y}; }
GDK_HD cplx csub(cplx a, cplx b) { return cplx{a.x - b.x, a.y - b.y}; }

struct CoopHost {
    int tid = 0, nt = 1;
    void sync() const {}
    double sum(double v) const { return v; }
    double max(double v) const { return v; }
    int any(int v) const { return v; }
    void sumv(double* v, int n) const {}
    // part[k] += sum_{y0 <= y <= y1} sum_{x0 <= x <= x1} wy[k*G + y] * A[y*G + x] * wx[k*G + x],  k < n
    void bilinear(const double* A, int G, int y0, int y1, int x0, int x1, const double* wx, const double* wy, int n,
                  double* part) const {
        for (int y = y0; y <= y1; y++)
            for (int x = x0; x <= x1; x++)
                for (int k = 0; k < n; k++) part[k] += (wy[k * G + y] * A[(size_t)y * G + x]) * wx[k * G + x];
    }
};

#if defined(__CUDACC__)
#define COOP_RING_STAGES 4
struct CoopBlock {
    int tid, nt;
    double* red;              // >= 32 doubles of shared memory
    double* redv = nullptr;   // optional: >= 8 * 32 doubles for the fused reduction of up to 8 values (sumv)
    double* ring = nullptr;   // optional: COOP_RING_STAGES stages of ring_rows x G doubles for `bilinear`
    int ring_rows = 0;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double t = 0;
        const int nw = (nt + 31) >> 5;
        for (int i = 0; i < nw; i++) t += red[i];  // same order in every thread
        return t;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double t = red[0];
        const int nw = (nt + 31) >> 5;
        for (int i = 1; i < nw; i++) t = fmax(t, red[i]);
        return t;
    }
    __device__ __forceinline__ int any(int v) const { return __syncthreads_or(v); }
    // part[k] += sum_{y0 <= y <= y1} sum_{x0 <= x <= x1} wy[k*G + y] * A[y*G + x] * wx[k*G + x],  k < n <= 6 (per-thread
    // partial sums).  A streams from L2 / HBM: the sweep is bound by the bytes in flight, so whole rows are copied
    // asynchronously (cp.async, 16 bytes per thread and piece) into a ring of shared-memory stages, COOP_RING_STAGES - 1
    // stages ahead of the one being consumed; a warp then takes one row of the stage, its lanes stride over x.
    __device__ __forceinline__ void bilinear(const double* __restrict__ A, int G, int y0, int y1, int x0, int x1,
                                             const double* wx, const double* wy, int n, double* part) const {
        const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
        if (!ring || (G & 1) || y1 < y0) {  // plain path: rows dealt to warps, four loads in flight per lane
            for (int y = y0 + wid; y <= y1; y += nw) {
                const double* row = A + (size_t)y * G;
                for (int xb = x0 + lane; xb <= x1; xb += 128) {
                    double v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) v[j] = xb + 32 * j <= x1 ? row[xb + 32 * j] : 0.0;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (xb + 32 * j <= x1)
                            for (int k = 0; k < n; k++) part[k] += (wy[k * G + y] * v[j]) * wx[k * G + xb + 32 * j];
                }
            }
            return;
        }
        const int R = ring_rows, S = COOP_RING_STAGES;
        const int nchunk = (y1 - y0 + R) / R;
        auto issue = [&](int c) {
            if (c < nchunk) {
                const int ya = y0 + c * R;
                const int rows = (y1 - ya + 1) < R ? (y1 - ya + 1) : R;
                const char* src = reinterpret_cast<const char*>(A + (size_t)ya * G);
                const unsigned dst = (unsigned)__cvta_generic_to_shared(ring + (size_t)(c % S) * R * G);
                const int n16 = rows * (G >> 1);
                for (int i = tid; i < n16; i += nt)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (unsigned)i), "l"(src + 16 * (size_t)i) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");  // one group per call, empty past the end
        };
        for (int c = 0; c < S - 1; c++) issue(c);
        for (int c = 0; c < nchunk; c++) {
            asm volatile("cp.async.wait_group %0;" ::"n"(COOP_RING_STAGES - 2) : "memory");  // this thread's pieces of chunk c
            __syncthreads();  // everybody's pieces of chunk c have landed, and everybody is done with chunk c - 1 ...
            issue(c + S - 1);  // ... whose stage is refilled now, S - 1 chunks ahead
            const int ya = y0 + c * R;
            const int rows = (y1 - ya + 1) < R ? (y1 - ya + 1) : R;
            const double* st = ring + (size_t)(c % S) * R * G;
            for (int rr = wid; rr < rows; rr += nw) {
                const int y = ya + rr;
                double wyk[6];
#pragma unroll
                for (int k = 0; k < 6; k++) wyk[k] = k < n ? wy[k * G + y] : 0.0;
                const double* row = st + (size_t)rr * G;
                for (int x = x0 + lane; x <= x1; x += 32) {
                    const double v = row[x];
#pragma unroll
                    for (int k = 0; k < 6; k++)
                        if (k < n) part[k] += (wyk[k] * v) * wx[k * G + x];
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // the ring is idle before the caller reuses shared memory
    }
    // sums of n <= 8 values at once (same order as `sum`): two barriers instead of 2 n
    __device__ __forceinline__ void sumv(double* v, int n) const {
        if (!redv) {
            for (int k = 0; k < n; k++) v[k] = sum(v[k]);
            return;
        }
        for (int k = 0; k < n; k++) {
            double t = v[k];
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            v[k] = t;
        }
        __syncthreads();
        if ((tid & 31) == 0)
            for (int k = 0; k < n; k++) redv[k * 32 + (tid >> 5)] = v[k];
        __syncthreads();
        const int nw = (nt + 31) >> 5;
        for (int k = 0; k < n; k++) {
            double t = 0;
            for (int i = 0; i < nw; i++) t += redv[k * 32 + i];
            v[k] = t;
        }
    }
};
#endif
