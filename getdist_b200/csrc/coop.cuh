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
};

#if defined(__CUDACC__)
struct CoopBlock {
    int tid, nt;
    double* red;              // >= 32 doubles of shared memory
    double* redv = nullptr;   // optional: >= 8 * 32 doubles for the fused reduction of up to 8 values (sumv)
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
