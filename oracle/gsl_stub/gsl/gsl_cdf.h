/* Stand-in for GSL's <gsl/gsl_cdf.h>, used ONLY to compile the unmodified reference
 * sources into oracle/_ref/ without building the vendored GSL tarball (≈5 min).
 *
 * TEST INFRASTRUCTURE -- not part of the product.
 *
 * The reference calls exactly one GSL function, gsl_cdf_chisq_P(x, nu), and only at
 * *print* time for SNP p-values (reference src/GenomeBwt.cpp:749,776,803,824), i.e.
 * downstream of the accumulators and outside the hot path.  The
 * definition below is the textbook regularized lower incomplete gamma P(nu/2, x/2)
 * (series for x < a+1, Lentz continued fraction otherwise).
 */
#ifndef GMX_ORACLE_GSL_CDF_STUB_H
#define GMX_ORACLE_GSL_CDF_STUB_H

#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

static inline double gmx_stub_gammp(double a, double x)
{
    if (!(x > 0.0)) return 0.0;
    double gln = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 100000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-16) break;
        }
        return sum * exp(-x + a * log(x) - gln);
    } else {
        const double tiny = 1e-300;
        double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
        for (int i = 1; i < 100000; ++i) {
            double an = -i * (i - a);
            b += 2.0;
            d = an * d + b; if (fabs(d) < tiny) d = tiny;
            c = b + an / c; if (fabs(c) < tiny) c = tiny;
            d = 1.0 / d;
            double del = d * c;
            h *= del;
            if (fabs(del - 1.0) < 1e-16) break;
        }
        return 1.0 - exp(-x + a * log(x) - gln) * h;
    }
}

static inline double gsl_cdf_chisq_P(double x, double nu)
{
    return gmx_stub_gammp(nu / 2.0, x / 2.0);
}

#ifdef __cplusplus
}
#endif

#endif
