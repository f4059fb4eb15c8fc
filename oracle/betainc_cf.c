/* Plain-C restatement of the regularised incomplete beta function I_x(a,b) that the
 * reference obtains from scipy.stats.beta.cdf -> scipy.special.betainc
 * (call sites src/normalisr/association.py:249, 379, 563, 717; scipy is unpinned in
 * setup.py:29, fixtures were generated with scipy 1.18.1).  TEST INFRASTRUCTURE ONLY:
 * nothing in normalisr_b200/ links or loads this file.
 *
 * Published algorithm restated here: the continued fraction DLMF 8.17.22
 *     I_x(a,b) = x^a (1-x)^b / (a B(a,b)) * 1/(1+ d1/(1+ d2/(1+ ...)))
 * evaluated with the modified Lentz recurrence, applied directly for
 * x < (a+1)/(a+b+2) and through I_x(a,b) = 1 - I_{1-x}(b,a) otherwise.  log B(a,b)
 * uses a Stirling difference for large arguments so that the prefactor keeps ~1e-15
 * relative accuracy at a ~ 5e5 (plain lgamma differences lose 6 digits there).
 * This is deliberately a different algorithm from the one the CUDA epilogue uses
 * (DiDonato-Morris asymptotic expansion), so agreement between the two is evidence.
 */
#include <math.h>
#include <stdint.h>

/* Stirling correction  lgamma(x) - [(x-1/2)ln x - x + ln(2pi)/2],  x >= 8 */
static double stirling_corr(double x) {
    double i = 1.0 / x, i2 = i * i;
    return i * (1.0 / 12 - i2 * (1.0 / 360 - i2 * (1.0 / 1260 - i2 * (1.0 / 1680 - i2 * (1.0 / 1188)))));
}

/* lgamma(a+b) - lgamma(a), accurate for large a */
static double lgamma_diff(double a, double b) {
    if (a < 8.0) return lgamma(a + b) - lgamma(a);
    return (a - 0.5) * log1p(b / a) + b * log(a + b) - b + (stirling_corr(a + b) - stirling_corr(a));
}

static double log_beta(double a, double b) {
    if (a < b) { double t = a; a = b; b = t; }
    return lgamma(b) - lgamma_diff(a, b);
}

static double betacf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 2000000; ++m) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}

double oracle_betainc(double a, double b, double x) {
    if (!(x > 0.0)) return 0.0;
    if (x >= 1.0) return 1.0;
    /* x^a (1-x)^b / B(a,b), in logs; log1p keeps 1-x accurate when x is near 1 */
    double lx = (x > 0.5) ? log1p(-(1.0 - x)) : log(x);
    double l1x = log1p(-x);
    double lbt = a * lx + b * l1x - log_beta(a, b);
    if (x < (a + 1.0) / (a + b + 2.0)) return exp(lbt) * betacf(a, b, x) / a;
    return 1.0 - exp(lbt) * betacf(b, a, 1.0 - x) / b;
}

/* out[i] = I_{x[i]}(a[i], b) */
void oracle_betainc_array(const double* x, const double* a, double b, double* out, int64_t n) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; ++i) out[i] = oracle_betainc(a[i], b, x[i]);
}
