// Principal eigenvector of a 3x3 covariance matrix with LAPACK dgeev's SIGN convention.
//
// greenhouse.py:229-230 does `w, v = np.linalg.eig(X_cov); d_l = v[:, np.argmax(w)]`.  The sign of
// d_l decides which bifurcation child is attached first, i.e. CSV row order, and
// LAPACK has no sign normalisation -- the sign is whatever falls out of
//   dgebal -> dgehrd (dgehd2) -> dorghr -> dhseqr (dlahqr) -> dtrevc3 -> unit 2-norm scaling.
// This header restates that pipeline for n = 3 (reference LAPACK 3.11 as bundled with
// OpenBLAS 0.3.30; written from the published algorithms, no code copied), so that the same
// discrete choices are made: Householder sign beta = -sign(alpha)*norm, the Francis double-shift
// sweep order, the Ahues-Tisseur deflation tests, dlanv2's standardisation rotation, dtrevc3's
// back-substitution with x(ki) = 1 and the final 1/||v||_2 scaling.  Rounding may differ from the
// Fortran build by a few ULP (BLAS-level association); the sign and the ordering do not.
//
// Supported input: finite, symmetric, no exactly-zero off-diagonal structure that would make
// dgebal permute (never the case for covariances of >= 2 jittered 3-D points).  If the Schur form
// keeps a 2x2 block (numerically repeated eigenvalues -> numpy returns complex dtype) `status`
// is set to 1 and the caller falls back to the documented convention.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define OCTA_EIG_HD __host__ __device__
#else
#define OCTA_EIG_HD
#endif

namespace octa {
namespace eig3 {

OCTA_EIG_HD inline double sign_of(double a, double b) { return (b >= 0.0 && !(b == 0.0 && signbit(b))) ? fabs(a) : -fabs(a); }

OCTA_EIG_HD inline double dlapy2(double x, double y) {
    const double xa = fabs(x), ya = fabs(y);
    const double w = xa > ya ? xa : ya, z = xa > ya ? ya : xa;
    if (z == 0.0) return w;
    const double q = z / w;
    return w * sqrt(1.0 + q * q);
}

// dnrm2 of n <= 3 values.  OpenBLAS' x86-64 dnrm2 accumulates in x87 extended precision and
// rounds sqrt once to double; on the host we do exactly that, on the device the sum of squares is
// formed error-free in double-double and the square root is corrected to the nearest double
// (agrees with the extended-precision result except for rare double-rounding ties).
OCTA_EIG_HD inline double dnrm2_small(int n, const double* x) {
    if (n < 1) return 0.0;
    if (n == 1) return fabs(x[0]);
#if defined(__CUDA_ARCH__)
    double hi = 0.0, lo = 0.0;
    for (int i = 0; i < n; ++i) {
        const double p = x[i] * x[i];
        const double pe = fma(x[i], x[i], -p);      // exact product = p + pe
        const double s = hi + p;                     // two-sum
        const double bb = s - hi;
        const double se = (hi - (s - bb)) + (p - bb);
        hi = s;
        lo += se + pe;
    }
    const double t = hi + lo;
    const double tl = lo - (t - hi);
    if (t == 0.0) return 0.0;
    const double r = sqrt(t);
    const double res = fma(-r, r, t) + tl;           // t + tl - r*r
    return r + res / (2.0 * r);
#else
    long double acc = 0.0L;
    for (int i = 0; i < n; ++i) acc += (long double)x[i] * (long double)x[i];
    return (double)sqrtl(acc);
#endif
}

// DROT as executed by OpenBLAS' x86-64 kernel: x' = fma(c, x, s*y), y' = fma(c, y, -(s*x))
OCTA_EIG_HD inline void drot1(double* x, double* y, double c, double s) {
    const double a = *x, b = *y;
    *x = fma(c, a, s * b);
    *y = fma(c, b, -(s * a));
}

// DLARFG: elementary reflector H = I - tau*[1;v][1;v]^T with H*[alpha;x] = [beta;0]
OCTA_EIG_HD inline void dlarfg(int n, double* alpha, double* x, double* tau) {
    if (n <= 1) { *tau = 0.0; return; }
    const double xnorm = dnrm2_small(n - 1, x);
    if (xnorm == 0.0) { *tau = 0.0; return; }
    const double beta = -sign_of(dlapy2(*alpha, xnorm), *alpha);
    // (the |beta| < safmin rescaling loop of DLARFG cannot trigger for covariance magnitudes)
    *tau = (beta - *alpha) / beta;
    const double sc = 1.0 / (*alpha - beta);
    for (int i = 0; i < n - 1; ++i) x[i] *= sc;
    *alpha = beta;
}

// DLANV2: Schur factorisation of a real 2x2 [[a b][c d]]; returns 0 if the block became upper
// triangular (real eigenvalues), 1 if a complex / numerically equal pair remains.
OCTA_EIG_HD inline int dlanv2(double* a, double* b, double* c, double* d, double* cs, double* sn) {
    const double eps = 2.220446049250313e-16;  // DLAMCH('P') = eps*base
    const double multpl = 4.0;
    if (*c == 0.0) {
        *cs = 1.0; *sn = 0.0;
    } else if (*b == 0.0) {
        *cs = 0.0; *sn = 1.0;
        const double temp = *d;
        *d = *a; *a = temp; *b = -*c; *c = 0.0;
    } else if ((*a - *d) == 0.0 && sign_of(1.0, *b) != sign_of(1.0, *c)) {
        *cs = 1.0; *sn = 0.0;
    } else {
        const double temp = *a - *d;
        double p = 0.5 * temp;
        const double bcmax = fmax(fabs(*b), fabs(*c));
        const double bcmis = fmin(fabs(*b), fabs(*c)) * sign_of(1.0, *b) * sign_of(1.0, *c);
        const double scale = fmax(fabs(p), bcmax);
        double z = (p / scale) * p + (bcmax / scale) * bcmis;
        if (z >= multpl * eps) {
            // real eigenvalues
            z = p + sign_of(sqrt(scale) * sqrt(z), p);
            *a = *d + z;
            *d = *d - (bcmax / z) * bcmis;
            const double tau = dlapy2(*c, z);
            *cs = z / tau;
            *sn = *c / tau;
            *b = *b - *c;
            *c = 0.0;
        } else {
            // complex eigenvalues, or real (almost) equal eigenvalues: make diagonal elements equal
            const double sigma = *b + *c;
            const double tau = dlapy2(sigma, temp);
            *cs = sqrt(0.5 * (1.0 + fabs(sigma) / tau));
            *sn = -(p / (tau * *cs)) * sign_of(1.0, sigma);
            const double aa = *a * *cs + *b * *sn, bb = -*a * *sn + *b * *cs;
            const double cc = *c * *cs + *d * *sn, dd = -*c * *sn + *d * *cs;
            *a = aa * *cs + cc * *sn;
            *b = bb * *cs + dd * *sn;
            *c = -aa * *sn + cc * *cs;
            *d = -bb * *sn + dd * *cs;
            const double t2 = 0.5 * (*a + *d);
            *a = t2; *d = t2;
            if (*c != 0.0) {
                if (*b != 0.0) {
                    if (sign_of(1.0, *b) == sign_of(1.0, *c)) {
                        // real eigenvalues: reduce to upper triangular form
                        const double sab = sqrt(fabs(*b)), sac = sqrt(fabs(*c));
                        p = sign_of(sab * sac, *c);
                        const double tau2 = 1.0 / sqrt(fabs(*b + *c));
                        *a = t2 + p;
                        *d = t2 - p;
                        *b = *b - *c;
                        *c = 0.0;
                        const double cs1 = sab * tau2, sn1 = sac * tau2;
                        const double t3 = *cs * cs1 - *sn * sn1;
                        *sn = *cs * sn1 + *sn * cs1;
                        *cs = t3;
                    }
                } else {
                    *b = -*c; *c = 0.0;
                    const double t3 = *cs;
                    *cs = -*sn; *sn = t3;
                }
            }
        }
    }
    return (*c != 0.0) ? 1 : 0;
}

#define H_(i, j) h[((i)-1) * 3 + ((j)-1)]
#define Z_(i, j) z[((i)-1) * 3 + ((j)-1)]

// DLAHQR for n = 3, ILO = 1, IHI = 3, WANTT = WANTZ = true.  h, z row-major 3x3.  Returns 0 on
// success, >0 if it failed to converge.
OCTA_EIG_HD inline int dlahqr3(double* h, double* z, double* trace = nullptr) {
    int ntr = 0;
    const double safmin = 2.2250738585072014e-308, ulp = 2.220446049250313e-16;
    const double dat1 = 0.75, dat2 = -0.4375;
    const int n = 3, ilo = 1, ihi = 3, kexsh = 10;
    H_(3, 1) = 0.0;
    const double smlnum = safmin * ((double)(ihi - ilo + 1) / ulp);
    const int i1 = 1, i2 = n;
    const int itmax = 30 * 10;
    int kdefl = 0;
    int i = ihi;
    while (true) {
        int l = ilo;
        if (i < ilo) break;
        bool converged = false;
        for (int its = 0; its <= itmax; ++its) {
            int k;
            for (k = i; k >= l + 1; --k) {
                if (fabs(H_(k, k - 1)) <= smlnum) break;
                double tst = fabs(H_(k - 1, k - 1)) + fabs(H_(k, k));
                if (tst == 0.0) {
                    if (k - 2 >= ilo) tst += fabs(H_(k - 1, k - 2));
                    if (k + 1 <= ihi) tst += fabs(H_(k + 1, k));
                }
                if (fabs(H_(k, k - 1)) <= ulp * tst) {
                    const double ab = fmax(fabs(H_(k, k - 1)), fabs(H_(k - 1, k)));
                    const double ba = fmin(fabs(H_(k, k - 1)), fabs(H_(k - 1, k)));
                    const double aa = fmax(fabs(H_(k, k)), fabs(H_(k - 1, k - 1) - H_(k, k)));
                    const double bb = fmin(fabs(H_(k, k)), fabs(H_(k - 1, k - 1) - H_(k, k)));
                    const double s = aa + ab;
                    if (ba * (ab / s) <= fmax(smlnum, ulp * (bb * (aa / s)))) break;
                }
            }
            l = k;
#ifdef OCTA_EIG_TRACE
            printf("its=%d i=%d l=%d  H21=%.6g H32=%.6g diag %.10g %.10g %.10g\n", its, i, l, H_(2,1), H_(3,2), H_(1,1), H_(2,2), H_(3,3));
#endif
            if (l > ilo) H_(l, l - 1) = 0.0;
            if (l >= i - 1) { converged = true; if (trace && ntr < 12) { trace[ntr++] = its; trace[ntr++] = l; trace[ntr++] = i; } break; }
            ++kdefl;
            double h11, h21, h12, h22;
            if (kdefl % (2 * kexsh) == 0) {
                const double s = fabs(H_(i, i - 1)) + fabs(H_(i - 1, i - 2));
                h11 = dat1 * s + H_(i, i); h12 = dat2 * s; h21 = s; h22 = h11;
            } else if (kdefl % kexsh == 0) {
                const double s = fabs(H_(l + 1, l)) + fabs(H_(l + 2, l + 1));
                h11 = dat1 * s + H_(l, l); h12 = dat2 * s; h21 = s; h22 = h11;
            } else {
                h11 = H_(i - 1, i - 1); h21 = H_(i, i - 1); h12 = H_(i - 1, i); h22 = H_(i, i);
            }
            double rt1r, rt1i, rt2r, rt2i;
            {
                const double s = fabs(h11) + fabs(h12) + fabs(h21) + fabs(h22);
                if (s == 0.0) {
                    rt1r = rt1i = rt2r = rt2i = 0.0;
                } else {
                    h11 /= s; h21 /= s; h12 /= s; h22 /= s;
                    const double tr = (h11 + h22) / 2.0;
                    const double det = (h11 - tr) * (h22 - tr) - h12 * h21;
                    const double rtdisc = sqrt(fabs(det));
                    if (det >= 0.0) {
                        rt1r = tr * s; rt2r = rt1r; rt1i = rtdisc * s; rt2i = -rt1i;
                    } else {
                        rt1r = tr + rtdisc; rt2r = tr - rtdisc;
                        if (fabs(rt1r - h22) <= fabs(rt2r - h22)) { rt1r = rt1r * s; rt2r = rt1r; }
                        else { rt2r = rt2r * s; rt1r = rt2r; }
                        rt1i = rt2i = 0.0;
                    }
                }
            }
            double v[3];
            int m;
            for (m = i - 2; m >= l; --m) {
                double h21s = fabs(H_(m + 1, m));
                double s = fabs(H_(m, m) - rt2r) + fabs(rt2i) + h21s;
                h21s = H_(m + 1, m) / s;
                v[0] = h21s * H_(m, m + 1) + (H_(m, m) - rt1r) * ((H_(m, m) - rt2r) / s) - rt1i * (rt2i / s);
                v[1] = h21s * (H_(m, m) + H_(m + 1, m + 1) - rt1r - rt2r);
                v[2] = h21s * H_(m + 2, m + 1);
                s = fabs(v[0]) + fabs(v[1]) + fabs(v[2]);
                v[0] /= s; v[1] /= s; v[2] /= s;
                if (m == l) break;
                const double h00 = fabs(H_(m - 1, m - 1)), h11a = fabs(H_(m, m)), h22a = fabs(H_(m + 1, m + 1));
                if (fabs(H_(m, m - 1)) * (fabs(v[1]) + fabs(v[2])) <= ulp * fabs(v[0]) * (h00 + h11a + h22a)) break;
            }
            if (m < l) m = l;
            for (int k2 = m; k2 <= i - 1; ++k2) {
                const int nr = (3 < i - k2 + 1) ? 3 : (i - k2 + 1);
                if (k2 > m) { for (int q = 0; q < nr; ++q) v[q] = H_(k2 + q, k2 - 1); }
                double t1;
                dlarfg(nr, &v[0], &v[1], &t1);
                if (k2 > m) {
                    H_(k2, k2 - 1) = v[0];
                    H_(k2 + 1, k2 - 1) = 0.0;
                    if (k2 < i - 1) H_(k2 + 2, k2 - 1) = 0.0;
                } else if (m > l) {
                    H_(k2, k2 - 1) = H_(k2, k2 - 1) * (1.0 - t1);
                }
                const double v2 = v[1], t2 = t1 * v2;
                if (nr == 3) {
                    const double v3 = v[2], t3 = t1 * v3;
                    for (int j = k2; j <= i2; ++j) {
                        const double sum = H_(k2, j) + v2 * H_(k2 + 1, j) + v3 * H_(k2 + 2, j);
                        H_(k2, j) -= sum * t1; H_(k2 + 1, j) -= sum * t2; H_(k2 + 2, j) -= sum * t3;
                    }
                    const int jmax = (k2 + 3 < i) ? k2 + 3 : i;
                    for (int j = i1; j <= jmax; ++j) {
                        const double sum = H_(j, k2) + v2 * H_(j, k2 + 1) + v3 * H_(j, k2 + 2);
                        H_(j, k2) -= sum * t1; H_(j, k2 + 1) -= sum * t2; H_(j, k2 + 2) -= sum * t3;
                    }
                    for (int j = 1; j <= 3; ++j) {
                        const double sum = Z_(j, k2) + v2 * Z_(j, k2 + 1) + v3 * Z_(j, k2 + 2);
                        Z_(j, k2) -= sum * t1; Z_(j, k2 + 1) -= sum * t2; Z_(j, k2 + 2) -= sum * t3;
                    }
                } else if (nr == 2) {
                    for (int j = k2; j <= i2; ++j) {
                        const double sum = H_(k2, j) + v2 * H_(k2 + 1, j);
                        H_(k2, j) -= sum * t1; H_(k2 + 1, j) -= sum * t2;
                    }
                    for (int j = i1; j <= i; ++j) {
                        const double sum = H_(j, k2) + v2 * H_(j, k2 + 1);
                        H_(j, k2) -= sum * t1; H_(j, k2 + 1) -= sum * t2;
                    }
                    for (int j = 1; j <= 3; ++j) {
                        const double sum = Z_(j, k2) + v2 * Z_(j, k2 + 1);
                        Z_(j, k2) -= sum * t1; Z_(j, k2 + 1) -= sum * t2;
                    }
                }
            }
        }
        if (!converged) return i;
        if (l == i - 1) {
            double cs, sn;
#ifdef OCTA_EIG_TRACE
            printf("dlanv2 in: a=%.17g b=%.17g c=%.17g d=%.17g\n", H_(i - 1, i - 1), H_(i - 1, i), H_(i, i - 1), H_(i, i));
#endif
            dlanv2(&H_(i - 1, i - 1), &H_(i - 1, i), &H_(i, i - 1), &H_(i, i), &cs, &sn);
#ifdef OCTA_EIG_TRACE
            printf("dlanv2 out: a=%.17g b=%.17g c=%.17g d=%.17g cs=%.17g sn=%.17g\n", H_(i - 1, i - 1), H_(i - 1, i), H_(i, i - 1), H_(i, i), cs, sn);
#endif
            // apply the rotation to the rest of H (DROT) and to Z
            for (int j = i + 1; j <= i2; ++j) drot1(&H_(i - 1, j), &H_(i, j), cs, sn);
            for (int j = i1; j <= i - 2; ++j) drot1(&H_(j, i - 1), &H_(j, i), cs, sn);
            for (int j = 1; j <= 3; ++j) drot1(&Z_(j, i - 1), &Z_(j, i), cs, sn);
        }
        kdefl = 0;
        i = l - 1;
    }
    return 0;
}

// Full pipeline.  cov: row-major symmetric 3x3.  out_w[3]: eigenvalues in LAPACK order;
// out_v: row-major 3x3, column k = unit eigenvector k.  Returns status: 0 ok, 1 = a 2x2 block is left
// (complex/equal pair: those two columns of out_v are zero, the real one is valid), 2 = no convergence.
OCTA_EIG_HD inline int dgeev3_sym(const double* cov, double* out_w, double* out_v, double* dbg = nullptr) {
    double h[9], z[9];
    for (int i = 0; i < 9; ++i) h[i] = cov[i];
    // DGEHD2, i = 1: reflector annihilating A(3,1)
    double alpha = H_(2, 1), x = H_(3, 1), tau;
    dlarfg(2, &alpha, &x, &tau);
    const double v2 = x;
    if (tau != 0.0) {
        // A(1:3,2:3) := A(1:3,2:3) * H   (DLARF 'Right': dgemv 'N' then dger, OpenBLAS association)
        double wv[3];
        for (int r = 1; r <= 3; ++r) wv[r - 1] = fma(H_(r, 3), v2, H_(r, 2));
        for (int r = 1; r <= 3; ++r) {
            H_(r, 2) = fma(wv[r - 1], -tau, H_(r, 2));
            H_(r, 3) = fma(wv[r - 1], (-tau) * v2, H_(r, 3));
        }
        // A(2:3,2:3) := H * A(2:3,2:3)  (DLARF 'Left': dgemv 'T' then dger)
        double wc[2];
        for (int c = 2; c <= 3; ++c) wc[c - 2] = H_(2, c) + H_(3, c) * v2;
        for (int c = 2; c <= 3; ++c) {
            const double t = (-tau) * wc[c - 2];
            H_(2, c) = H_(2, c) + t;
            H_(3, c) = fma(v2, t, H_(3, c));
        }
    }
    H_(2, 1) = alpha;
    H_(3, 1) = 0.0;
    // DORGHR
    Z_(1, 1) = 1.0; Z_(1, 2) = 0.0; Z_(1, 3) = 0.0; Z_(2, 1) = 0.0; Z_(3, 1) = 0.0;
    Z_(2, 2) = 1.0 - tau; Z_(3, 2) = -tau * v2;
    Z_(2, 3) = (-tau) * v2; Z_(3, 3) = fma(v2, (-tau) * v2, 1.0);
    if (dbg) { for (int q = 0; q < 9; ++q) { dbg[q] = h[q]; dbg[9 + q] = z[q]; } }
    if (dlahqr3(h, z, dbg ? dbg + 36 : nullptr) != 0) return 2;
    if (dbg) { for (int q = 0; q < 9; ++q) { dbg[18 + q] = h[q]; dbg[27 + q] = z[q]; } }
    // eigenvalues (real parts) in LAPACK order; a surviving 2x2 block is a complex / numerically
    // equal pair (numpy then returns complex dtype and the reference keeps np.real of the result)
    const bool blk12 = H_(2, 1) != 0.0, blk23 = H_(3, 2) != 0.0;
    for (int k = 0; k < 3; ++k) out_w[k] = h[4 * k];
    // DTREVC3 (real eigenvectors, back-transformed with Z) followed by dgeev's 1/||v||_2 scaling
    const double ulp = 2.220446049250313e-16, unfl = 2.2250738585072014e-308;
    const double smlnum = unfl * (3.0 / ulp);
    int status = (blk12 || blk23) ? 1 : 0;
    for (int ki = 3; ki >= 1; --ki) {
        const bool in_block = (blk12 && ki <= 2) || (blk23 && ki >= 2);
        if (in_block) {   // complex pair member: no real eigenvector; leave zeros
            for (int r = 0; r < 3; ++r) out_v[r * 3 + (ki - 1)] = 0.0;
            continue;
        }
        const double wr = H_(ki, ki);
        const double smin = fmax(ulp * fabs(wr), smlnum);
        double xv[3] = {0.0, 0.0, 0.0};
        xv[ki - 1] = 1.0;
        for (int k = 1; k <= ki - 1; ++k) xv[k - 1] = -H_(k, ki);
        if (ki == 3 && blk12) {
            // 2x2 diagonal block (DLALN2, na = 2, nw = 1): (T(1:2,1:2) - wr I) x = b, complete pivoting
            double a11 = H_(1, 1) - wr, a12 = H_(1, 2), a21 = H_(2, 1), a22 = H_(2, 2) - wr;
            double b1 = xv[0], b2 = xv[1];
            const double m11 = fabs(a11), m12 = fabs(a12), m21 = fabs(a21), m22 = fabs(a22);
            double cmax = fmax(fmax(m11, m12), fmax(m21, m22));
            if (cmax < smin) { xv[0] = b1 / smin; xv[1] = b2 / smin; }
            else {
                // bring the pivot to (1,1)
                bool swap_rows = false, swap_cols = false;
                if (cmax == m11) {}
                else if (cmax == m12) swap_cols = true;
                else if (cmax == m21) swap_rows = true;
                else { swap_rows = true; swap_cols = true; }
                if (swap_rows) { double t; t = a11; a11 = a21; a21 = t; t = a12; a12 = a22; a22 = t; t = b1; b1 = b2; b2 = t; }
                if (swap_cols) { double t; t = a11; a11 = a12; a12 = t; t = a21; a21 = a22; a22 = t; }
                const double ur11r = 1.0 / a11, lr21 = ur11r * a21;
                double ur22 = a22 - a12 * lr21;
                if (fabs(ur22) < smin) ur22 = smin;
                const double br2 = b2 - lr21 * b1;
                const double x2 = br2 / ur22;
                const double x1 = b1 * ur11r - x2 * (ur11r * a12);
                if (swap_cols) { xv[0] = x2; xv[1] = x1; } else { xv[0] = x1; xv[1] = x2; }
            }
        } else {
            for (int j = ki - 1; j >= 1; --j) {
                // DLALN2, 1x1 real: (T(j,j) - wr) * X = B   (scale = 1: |B| <= ||T||, |csr| >= smin)
                double csr = H_(j, j) - wr;
                if (fabs(csr) < smin) csr = smin;
                const double xx = xv[j - 1] / csr;
                xv[j - 1] = xx;
                for (int k = 1; k <= j - 1; ++k) xv[k - 1] += -xx * H_(k, j);
            }
        }
        // back-transform: VR(:,ki) = Z(:,1:ki-1)*x(1:ki-1) + x(ki)*Z(:,ki)
        double vec[3];
        for (int r = 1; r <= 3; ++r) {
            double acc = xv[ki - 1] * Z_(r, ki);
            for (int k = 1; k <= ki - 1; ++k) acc += Z_(r, k) * xv[k - 1];
            vec[r - 1] = acc;
        }
        const double emax = fmax(fabs(vec[0]), fmax(fabs(vec[1]), fabs(vec[2])));
        const double remax = 1.0 / emax;
        for (int r = 0; r < 3; ++r) vec[r] *= remax;
        const double scl = 1.0 / dnrm2_small(3, vec);
        for (int r = 0; r < 3; ++r) out_v[r * 3 + (ki - 1)] = vec[r] * scl;
    }
    return status;
}

// greenhouse.py:229-233: d_l = real part of the eigenvector of argmax(w).  Returns 0 when the
// principal eigenvalue is real (always, up to a ~1e-8 corner: two numerically equal LARGEST
// eigenvalues, status 3) and 2 on non-convergence.
OCTA_EIG_HD inline int principal_axis(const double* cov, double* dl) {
    double w[3], v[9];
    const int st = dgeev3_sym(cov, w, v);
    if (st == 2) return 2;
    int k = 0;
    for (int i = 1; i < 3; ++i) if (w[i] > w[k]) k = i;
    dl[0] = v[k]; dl[1] = v[3 + k]; dl[2] = v[6 + k];
    if (st == 1 && dl[0] == 0.0 && dl[1] == 0.0 && dl[2] == 0.0) return 3;
    return 0;
}

#undef H_
#undef Z_

}  // namespace eig3
}  // namespace octa
