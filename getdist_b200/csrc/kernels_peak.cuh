// kernels_peak.cuh -- in-tree microbenchmarks behind gdk_measure_peaks(): the denominators of the roofline figures
// bench.py reports for kernels that are NOT bound by HBM bandwidth ("conv FLOP/s vs sm_100a peak",
// "achieved updates/s").  Each kernel is a few milliseconds of the one instruction mix it measures:
//   k_peak_fp64      independent DFMA chains                      -> FP64 FMA TFLOP/s (k_conv2d, k_cov_tiles, k_bw2d)
//   k_peak_atoms     64-bit fixed-point add into two 32-bit shared limbs (ATOMS with return + carry + RED), the update
//                    every histogram kernel issues; lanes on distinct banks (conflict free) or on random bins of a
//                    96 x 96 window (what a privatised 2D histogram sees)   -> updates/s
//   k_peak_l2red     REDG.ADD.64 at random addresses of an L2-resident region  -> reductions/s (window misses, flushes,
//                    k_hist2d_tiles)
//   k_peak_read      16-byte streaming loads, sum in registers    -> HBM read GB/s (a read-only sweep, which is what the
//                    histogram / moment / quantile passes are; MEASURED_PEAKS.json holds the driver's copy figure)
#pragma once
#include <stdint.h>

#include "kernels_quant.cuh"

__global__ void __launch_bounds__(512) k_peak_fp64(double* __restrict__ out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = a + threadIdx.x * 1e-9 + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = fma(x[k], b, a);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

// mode 0: lane -> its own bank (conflict free); mode 1: pseudo-random bin of a 96 x 96 window per update
__global__ void __launch_bounds__(512, 1) k_peak_atoms(unsigned long long* __restrict__ out, int iters, int mode) {
    extern __shared__ unsigned psm[];  // lo[NB], hi[NB]
    constexpr unsigned NB = 96 * 96;
    for (int i = threadIdx.x; i < 2 * NB; i += blockDim.x) psm[i] = 0;
    __syncthreads();
    unsigned base = (unsigned)__cvta_generic_to_shared(psm);
    asm volatile("mov.u32 %0, %0;" : "+r"(base));
    unsigned r = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    const unsigned long long w = 0x0000002000000001ull + threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned bin;
            if (mode == 0) {
                bin = (threadIdx.x & 31) + 32u * ((threadIdx.x >> 5) + 16u * k);  // warp-contiguous: one bank per lane
            } else {
                r = r * 1664525u + 1013904223u;
                bin = (r >> 8) % NB;
            }
            smem_add_u64_addr(base + (bin << 2), NB * 4u, w);
        }
    }
    __syncthreads();
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < (int)NB; i += blockDim.x) s += ((unsigned long long)psm[NB + i] << 32) | psm[i];
    if (s == 1) out[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak_l2red(unsigned long long* __restrict__ bins, unsigned nbins_mask, int iters) {
    unsigned r = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 977u;
    for (int i = 0; i < iters; i++) {
        r = r * 1664525u + 1013904223u;
        atomicAdd(bins + ((r >> 6) & nbins_mask), 1ull);
    }
}

__global__ void __launch_bounds__(512) k_peak_read(const double* __restrict__ x, int64_t n2, double* __restrict__ out) {
    double s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x * 4) {
        double2 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t j = i + (int64_t)k * gridDim.x * blockDim.x;
            v[k] = j < n2 ? ldg_stream2(x + 2 * j) : double2{0, 0};
        }
#pragma unroll
        for (int k = 0; k < 4; k++) s += v[k].x + v[k].y;
    }
    if (s == 12345.678) out[0] = s;
}
