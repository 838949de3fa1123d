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
};

#if defined(__CUDACC__)
struct CoopBlock {
    int tid, nt;
    double* red;  // >= 32 doubles of shared memory
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
};
#endif
