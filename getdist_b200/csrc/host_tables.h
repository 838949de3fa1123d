// host_tables.h -- host-side construction of the constant tables the grid-stage kernels read
// (twiddle factors, cosine tables, ISJ constants).  Built with libm on the host so that they carry
// the same correctly-rounded values numpy/scipy use.
#pragma once
#include <math.h>

#include "coop.cuh"
#include "kde1d_core.cuh"

// out[k] = exp(-2 pi i k / n) for k < count, with exact values on the axes
static inline void gdk_fill_roots(cplx* out, int n, int count) {
    for (int k = 0; k < count; k++) {
        double c, s;
        // exact quadrant symmetries when 4 | n (keeps 1, -i, -1, i exact and the table symmetric)
        if (n % 4 == 0) {
            const int q = n / 4;
            const int quad = (k / q) % 4, rem = k % q;
            const double ang = 2.0 * M_PI * (double)rem / (double)n;
            double c0 = cos(ang), s0 = sin(ang);
            if (rem == 0) {
                c0 = 1.0;
                s0 = 0.0;
            }
            switch (quad) {
                case 0: c = c0; s = s0; break;
                case 1: c = -s0; s = c0; break;
                case 2: c = -c0; s = -s0; break;
                default: c = s0; s = -c0; break;
            }
        } else {
            const double ang = 2.0 * M_PI * (double)k / (double)n;
            c = cos(ang);
            s = sin(ang);
        }
        out[k] = cplx{c, -s};
    }
}

// out[j] = cos(2 pi j / n4), j < n4 (n4 = 4n)
static inline void gdk_fill_cos(double* out, int n4) {
    const int q = n4 / 4;
    for (int j = 0; j < n4; j++) {
        const int quad = j / q, rem = j % q;
        const double ang = 2.0 * M_PI * (double)rem / (double)n4;
        double c0 = rem == 0 ? 1.0 : cos(ang), s0 = rem == 0 ? 0.0 : sin(ang);
        double c;
        switch (quad) {
            case 0: c = c0; break;
            case 1: c = -s0; break;
            case 2: c = -c0; break;
            default: c = s0; break;
        }
        out[j] = c;
    }
}

// kde_bandwidth.py:47-56
static inline void gdk_fill_isj_consts(IsjConsts* K) {
    const double rootpi = sqrt(M_PI);
    K->rootpi = rootpi;
    K->pi2 = M_PI * M_PI;
    for (int j = 0; j < 8; j++) {
        K->two_pi_pow[j] = 2 * pow(M_PI, 2 * j);
        K->cj[j] = 0;
    }
    for (int j = 6; j >= 2; j--) {
        double prod = 1;
        for (int o = 1; o < 2 * j; o += 2) prod *= o;
        K->cj[j] = (1 + pow(0.5, j + 0.5)) / 3 * prod / (rootpi / sqrt(2.0));
    }
}

#include "kde2d_core.cuh"
// kde_bandwidth.py:140-143 and the powers of pi the psi functionals use
static inline void gdk_fill_kde2d_consts(Kde2dConsts* K) {
    K->pi2 = M_PI * M_PI;
    K->K[0] = 1 / sqrt(2 * M_PI);
    for (int j = 1; j < 5; j++) {
        double prod = 1;
        for (int o = 1; o < 2 * j; o += 2) prod *= o;
        K->K[j] = ((j & 1) ? -1.0 : 1.0) * prod / sqrt(2 * M_PI);
    }
    K->Kodd[0] = 1;
    for (int j = 1; j < 9; j++) {
        double prod = 1;
        for (int o = 1; o < 2 * j; o += 2) prod *= o;
        K->Kodd[j] = prod / pow(2.0, j + 1) / sqrt(M_PI);
    }
    for (int k = 0; k < 12; k++) {
        K->pipow[k] = pow(M_PI, 2 * k);
        K->twopipow[k] = pow(2 * M_PI, k);
    }
}
