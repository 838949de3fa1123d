// solvers.cuh -- scalar root finders evaluated cooperatively (every thread of the group runs the same
// control flow on identical values; the function itself is a group-wide reduction).
//
// The reference calls scipy.optimize.fsolve (MINPACK hybrd, n = 1; kde_bandwidth.py:123) and
// scipy.optimize.brentq (kde_bandwidth.py:127, 162).  scipy ships no source in this image, so both are
// restated from the published algorithms (MINPACK User Guide / hybrd.f; Brent 1973 as arranged in scipy's
// Zeros/brentq.c) and validated against the live scipy calls in tests/.
//   - brentq_port follows scipy's iteration step for step: t* of the 2D optimiser is only converged to
//     ~1e-3 relative, so the sequence of evaluation points must be the same.
//   - hybrd1_port reproduces hybrd's trust-region/dogleg/Broyden logic specialised to one unknown,
//     with fsolve's defaults (epsfcn = machine eps, mode 1 scaling, maxfev = 400).
#pragma once
#include "coop.cuh"

#define GDK_DBL_EPS 2.220446049250313e-16

struct RootResult {
    double x;
    int nfev;
    int status;  // 0 ok; 1 function failure (non-finite / zero f); 2 sign error; 3 no convergence
};

// F: double operator()(double x, int& fail)
template <class F>
GDK_HD RootResult brentq_port(F& f, double xa, double xb, double xtol, double rtol, int maxiter) {
    RootResult r{0.0, 0, 0};
    double xpre = xa, xcur = xb, xblk = 0., fblk = 0., spre = 0., scur = 0.;
    int fail = 0;
    double fpre = f(xpre, fail);
    r.nfev++;
    if (fail) {
        r.status = 1;
        return r;
    }
    double fcur = f(xcur, fail);
    r.nfev++;
    if (fail) {
        r.status = 1;
        return r;
    }
    if (fpre == 0) {
        r.x = xpre;
        return r;
    }
    if (fcur == 0) {
        r.x = xcur;
        return r;
    }
    if ((fpre < 0) == (fcur < 0) && !(fpre != fpre) && !(fcur != fcur)) {
        r.status = 2;  // f(a) f(b) > 0  -> ValueError in scipy
        return r;
    }
    if (fpre != fpre || fcur != fcur) {
        r.status = 1;
        return r;
    }
    for (int i = 0; i < maxiter; i++) {
        if (fpre != 0 && fcur != 0 && ((fpre < 0) != (fcur < 0))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur;
            xcur = xblk;
            xblk = xpre;
            fpre = fcur;
            fcur = fblk;
            fblk = fpre;
        }
        const double delta = (xtol + rtol * fabs(xcur)) / 2;
        const double sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) {
            r.x = xcur;
            return r;
        }
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);  // secant
            } else {
                const double dpre = (fpre - fcur) / (xpre - xcur);
                const double dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));  // inverse quadratic
            }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) {
                spre = scur;
                scur = stry;
            } else {
                spre = sbis;
                scur = sbis;
            }
        } else {
            spre = sbis;
            scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta)
            xcur += scur;
        else
            xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur, fail);
        r.nfev++;
        if (fail || fcur != fcur) {
            r.status = 1;
            return r;
        }
    }
    r.status = 3;
    r.x = xcur;
    return r;
}

// MINPACK hybrd specialised to n = 1 with scipy.fsolve defaults (mode 1, epsfcn -> eps, nprint 0).
template <class F>
GDK_HD RootResult hybrd1_port(F& f, double x0, double xtol, double factor, int maxfev) {
    const double p1 = 0.1, p5 = 0.5, p001 = 1e-3, p0001 = 1e-4;
    const double epsmch = GDK_DBL_EPS;
    RootResult r{x0, 0, 0};
    int fail = 0;
    double x = x0;
    double fvec = f(x, fail);
    r.nfev = 1;
    if (fail) {
        r.status = 1;
        return r;
    }
    double fnorm = fabs(fvec);
    int iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0;
    double diag = 1, delta = 0, xnorm = 0;
    int info = 0;
    const double eps_fd = sqrt(epsmch);  // sqrt(max(epsfcn, epsmch))
    while (true) {  // outer loop: (re)compute the jacobian by forward differences
        bool jeval = true;
        double hstep = eps_fd * fabs(x);
        if (hstep == 0) hstep = eps_fd;
        double fp = f(x + hstep, fail);
        r.nfev++;
        if (fail) {
            r.status = 1;
            return r;
        }
        double J = (fp - fvec) / hstep;
        // qr factorisation of a 1x1 matrix: |r| = |J| (column norm), q = +-1.  Working directly with
        // J and fvec is equivalent (the sign cancels in every product hybrd forms).
        const double acnorm = fabs(J);
        if (iter == 1) {
            diag = acnorm;
            if (diag == 0) diag = 1;
            xnorm = fabs(diag * x);
            delta = factor * xnorm;
            if (delta == 0) delta = factor;
        }
        double qtf = fvec;  // (q^T f) up to the common sign
        double R = J;
        diag = fmax(diag, acnorm);
        while (true) {  // inner loop
            // dogleg, n = 1: Gauss-Newton step clipped to the trust region in scaled units
            double Rd = R;
            if (Rd == 0) {
                double t = epsmch * fabs(R);
                if (t == 0) t = epsmch;
                Rd = t;
            }
            double gn = qtf / Rd;  // dogleg's x (hybrd negates it afterwards)
            const double qnorm = fabs(diag * gn);
            double step;
            if (qnorm <= delta) {
                step = gn;
            } else {
                // scaled gradient direction is collinear with gn when n = 1; its minimiser along the
                // direction has norm sgnorm == qnorm > delta, so the step is delta along that direction
                const double g = R * qtf / diag;
                if (g == 0) {
                    step = (delta / qnorm) * gn;  // alpha = delta/qnorm branch of dogleg
                } else {
                    step = delta * (g > 0 ? 1.0 : -1.0) / diag;
                }
            }
            const double p = -step;
            const double xnew = x + p;
            const double pnorm = fabs(diag * p);
            if (iter == 1) delta = fmin(delta, pnorm);
            const double fnew = f(xnew, fail);
            r.nfev++;
            if (fail) {
                r.status = 1;
                return r;
            }
            const double fnorm1 = fabs(fnew);
            double actred = -1;
            if (fnorm1 < fnorm) actred = 1 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            const double w3 = R * p + qtf;  // predicted residual
            const double temp = fabs(w3);
            double prered = 0;
            if (temp < fnorm) prered = 1 - (temp / fnorm) * (temp / fnorm);
            double ratio = 0;
            if (prered > 0) ratio = actred / prered;
            if (ratio < p1) {
                ncsuc = 0;
                ncfail++;
                delta = p5 * delta;
            } else {
                ncfail = 0;
                ncsuc++;
                if (ratio >= p5 || ncsuc > 1) delta = fmax(delta, pnorm / p5);
                if (fabs(ratio - 1) <= p1) delta = pnorm / p5;
            }
            if (ratio >= p0001) {
                x = xnew;
                fvec = fnew;
                xnorm = fabs(diag * x);
                fnorm = fnorm1;
                iter++;
            }
            nslow1++;
            if (actred >= p001) nslow1 = 0;
            if (jeval) nslow2++;
            if (actred >= p1) nslow2 = 0;
            if (delta <= xtol * xnorm || fnorm == 0) info = 1;
            if (info != 0) break;
            if (r.nfev >= maxfev) info = 2;
            if (p1 * fmax(p1 * delta, pnorm) <= epsmch * xnorm) info = 3;
            if (nslow2 == 5) info = 4;
            if (nslow1 == 10) info = 5;
            if (info != 0) break;
            if (ncfail == 2) break;  // recompute the jacobian
            // Broyden rank-one update; for n = 1 it is the secant slope through x_old and x_old + p
            // (sum = q^T f(x+p);  wa2 = (sum - w3)/pnorm;  r += wa2 * diag*(diag*p)/pnorm)
            const double sum = fnew;
            const double wa2 = (sum - w3) / pnorm;
            const double wa1 = diag * ((diag * p) / pnorm);
            R = R + wa2 * wa1;
            if (ratio >= p0001) qtf = sum;
            jeval = false;
        }
        if (info != 0) break;
    }
    r.x = x;
    r.status = 0;  // fsolve returns x whatever info is (warnings are suppressed by the reference)
    return r;
}
