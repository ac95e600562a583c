/*
 * oracle/psi_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the only native code on the reference's hot path:
 * the C++ loop that `weave.inline` compiles at run time inside
 * geepee/kernels.py:181-240 (`compute_psi_weave`, loop body at 201-234).
 * The reference needs scipy.weave + blitz converters (absent here), so the
 * loop is restated with flat row-major indexing instead of blitz accessors.
 * Same loop order (n, m1, m2<=m1, q), same expression order, IEEE double,
 * libm exp -- so results agree with a weave build to the last bits libm
 * allows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path is CUDA only.
 *
 * Build: see oracle/Makefile  ->  oracle/_ref/libgeepee_oracle.so
 */
#include <math.h>
#include <stddef.h>

/* kernels.py:181-240.  Inputs exactly as the reference passes them to C:
 *   ls2[Q]   = exp(lls2)            (squared lengthscales, kernels.py:182)
 *   sf2      = exp(lsf2)            (signal variance,      kernels.py:183)
 *   log_denom_psi1[N*Q] = 0.5*log(ls2/(ls2+xvar))      (kernels.py:191-193)
 *   log_denom_psi2[N*Q] = 0.5*log(ls2/(ls2+2*xvar))    (kernels.py:188-190)
 * Outputs psi1[N*M], psi2[N*M*M] (row-major).
 */
void geepee_oracle_psi(long N, long M, long Q, double sf2,
                       const double *ls2, const double *z,
                       const double *xmean, const double *xvar,
                       const double *log_denom_psi1,
                       const double *log_denom_psi2,
                       double *psi1, double *psi2)
{
    for (long n = 0; n < N; n++) {
        const double *mu = xmean + n * Q;
        const double *vv = xvar + n * Q;
        const double *ld1 = log_denom_psi1 + n * Q;
        const double *ld2 = log_denom_psi2 + n * Q;
        double *p2 = psi2 + (size_t)n * M * M;
        for (long m1 = 0; m1 < M; m1++) {
            double log_psi1 = 0;
            for (long m2 = 0; m2 <= m1; m2++) {
                double log_psi2 = 0;
                for (long q = 0; q < Q; q++) {
                    double vq = vv[q];
                    double lq = ls2[q];
                    double z1q = z[m1 * Q + q];
                    double z2q = z[m2 * Q + q];
                    if (m2 == 0) {
                        double muz = mu[q] - z1q;
                        log_psi1 += -muz * muz / 2.0 / (vq + lq) + ld1[q];
                    }
                    double muzhat = mu[q] - (z1q + z2q) / 2.0;
                    double dz = z1q - z2q;
                    log_psi2 += -dz * dz / (4.0 * lq)
                                - muzhat * muzhat / (2.0 * vq + lq) + ld2[q];
                }
                double e = exp(log_psi2);
                p2[m1 * M + m2] = sf2 * sf2 * e;
                if (m1 != m2)
                    p2[m2 * M + m1] = sf2 * sf2 * e;
            }
            psi1[n * M + m1] = sf2 * exp(log_psi1);
        }
    }
}

/* kernels.py:10-22 (`compute_kernel`): scipy cdist(...,'seuclidean',V=ls)
 * squared, i.e. r2 = sum_q (x_q - z_q)^2 / ls2_q ; k = sf2*exp(-r2/2).
 * Restated because scipy's cdist is the third-party piece of that function. */
void geepee_oracle_kernel(long N, long M, long Q, double sf2,
                          const double *ls2, const double *x, const double *z,
                          double *k)
{
    for (long n = 0; n < N; n++)
        for (long m = 0; m < M; m++) {
            double r2 = 0;
            for (long q = 0; q < Q; q++) {
                double d = x[n * Q + q] - z[m * Q + q];
                r2 += d * d / ls2[q];
            }
            k[n * M + m] = sf2 * exp(-0.5 * r2);
        }
}
