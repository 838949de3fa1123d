// fft.cuh -- batched Stockham autosort FFT (radix-4 stages + one radix-2 stage when log2 n is odd)
// and the DCT-II built on it, for power-of-two line lengths; direct O(n^2) table-driven DFT/DCT for
// other lengths.  Lines live in shared memory on the device (ping-pong buffers a/b).
//
// Replaces, on the reference path: scipy.fftpack.dct (kde_bandwidth.py:116, convolve.py:565-566) and
// np.fft.fft2 (kde_bandwidth.py:156).  Conventions:
//   DFT   X_k = sum_i x_i exp(-2 pi i ik/n)
//   DCT-II y_k = 2 sum_i x_i cos(pi k (2i+1) / (2n))          (scipy.fftpack.dct, type 2, norm=None)
#pragma once
#include "coop.cuh"

GDK_HD bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// Forward FFT of nl lines of length n (power of two, n >= 2).  Line l occupies a[l*n .. l*n+n).
// tw[k] = exp(-2 pi i k / n), k < n.  Returns the buffer (a or b) that holds the result.
template <class C>
GDK_HD cplx* fft_lines(const C& co, cplx* a, cplx* b, int n, int nl, const cplx* tw) {
    cplx* x = a;
    cplx* y = b;
    int m = n, s = 1;
    for (; m >= 4; m >>= 2, s <<= 2) {
        const int n1 = m >> 2;
        const int tws = n / m;  // twiddle stride: exp(-2 pi i p/m) = tw[p * n/m]
        const int per_line = n >> 2;
        const int total = nl * per_line;
        for (int it = co.tid; it < total; it += co.nt) {
            const int l = it / per_line;
            const int idx = it - l * per_line;
            const int p = idx / s;
            const int q = idx - p * s;
            const cplx* xl = x + (size_t)l * n;
            cplx* yl = y + (size_t)l * n;
            const cplx A = xl[q + s * p];
            const cplx B = xl[q + s * (p + n1)];
            const cplx Cc = xl[q + s * (p + 2 * n1)];
            const cplx D = xl[q + s * (p + 3 * n1)];
            const cplx apc = cadd(A, Cc), amc = csub(A, Cc), bpd = cadd(B, D), bmd = csub(B, D);
            const cplx jbmd = cplx{-bmd.y, bmd.x};  // i * (b - d)
            const cplx w1 = tw[p * tws], w2 = tw[2 * p * tws], w3 = tw[3 * p * tws];
            yl[q + s * (4 * p)] = cadd(apc, bpd);
            yl[q + s * (4 * p + 1)] = cmul(w1, csub(amc, jbmd));
            yl[q + s * (4 * p + 2)] = cmul(w2, csub(apc, bpd));
            yl[q + s * (4 * p + 3)] = cmul(w3, cadd(amc, jbmd));
        }
        co.sync();
        cplx* t = x;
        x = y;
        y = t;
    }
    if (m == 2) {  // final radix-2 stage: s == n/2, p == 0, twiddle 1
        const int per_line = n >> 1;
        const int total = nl * per_line;
        for (int it = co.tid; it < total; it += co.nt) {
            const int l = it / per_line;
            const int q = it - l * per_line;
            const cplx* xl = x + (size_t)l * n;
            cplx* yl = y + (size_t)l * n;
            const cplx A = xl[q], B = xl[q + s];
            yl[q] = cadd(A, B);
            yl[q + s] = csub(A, B);
        }
        co.sync();
        cplx* t = x;
        x = y;
        y = t;
    }
    return x;
}

// DCT-II of nl real lines of length n (power of two).  in/out: real, line stride n (may alias).
// a, b: complex scratch of nl*n each.  tw: n-th roots (as above); tw4[k] = exp(-2 pi i k/(4n)), k < n.
template <class C>
GDK_HD void dct2_lines_pow2(const C& co, const double* in, double* out, cplx* a, cplx* b, int n, int nl,
                            const cplx* tw, const cplx* tw4) {
    const int h = n >> 1;
    for (int it = co.tid; it < nl * h; it += co.nt) {
        const int l = it / h, i = it - l * h;
        a[(size_t)l * n + i] = cplx{in[(size_t)l * n + 2 * i], 0.0};
        a[(size_t)l * n + n - 1 - i] = cplx{in[(size_t)l * n + 2 * i + 1], 0.0};
    }
    co.sync();
    const cplx* V = fft_lines(co, a, b, n, nl, tw);
    for (int it = co.tid; it < nl * n; it += co.nt) {
        const int k = it % n;
        const cplx v = V[it];
        out[it] = 2.0 * (v.x * tw4[k].x - v.y * tw4[k].y);
    }
    co.sync();
}

// Direct DCT-II for any n: cos4[j] = cos(2 pi j / (4n)), j < 4n.  in and out must not alias.
template <class C>
GDK_HD void dct2_lines_direct(const C& co, const double* in, double* out, int n, int nl, const double* cos4) {
    const int n4 = 4 * n;
    for (int it = co.tid; it < nl * n; it += co.nt) {
        const int l = it / n, k = it - l * n;
        const double* x = in + (size_t)l * n;
        double acc = 0;
        int j = k % n4;            // k*(2i+1) mod 4n, advanced incrementally by 2k
        const int step = (2 * k) % n4;
        for (int i = 0; i < n; i++) {
            acc += x[i] * cos4[j];
            j += step;
            if (j >= n4) j -= n4;
        }
        out[it] = 2.0 * acc;
    }
    co.sync();
}

// Direct DFT for any n: tw[j] = exp(-2 pi i j/n).  in and out must not alias.
template <class C>
GDK_HD void dft_lines_direct(const C& co, const cplx* in, cplx* out, int n, int nl, const cplx* tw) {
    for (int it = co.tid; it < nl * n; it += co.nt) {
        const int l = it / n, k = it - l * n;
        const cplx* x = in + (size_t)l * n;
        cplx acc{0, 0};
        int j = 0;
        for (int i = 0; i < n; i++) {
            acc = cadd(acc, cmul(x[i], tw[j]));
            j += k;
            if (j >= n) j -= n;
        }
        out[it] = acc;
    }
    co.sync();
}
