// Exact two-sided P-value of a partial correlation:  P = I_{1-r^2}(a, 1/2),
// a = (n - 1 - rank(C) - dimreduce) / 2.
// Replaces scipy.stats.beta.cdf(1 - R2, a, 0.5) at reference
// src/normalisr/association.py:249 and :563.
//
// Two evaluation routes, both in float64 and both usable on host and device
// (the host build is what tests/test_pvalue_host.py checks against scipy/mpmath):
//   * a >= 15 and r^2 < 0.3  (every pair of a real data set): DiDonato & Morris'
//     large-a asymptotic expansion (ACM TOMS 18 (1992) 360, "BGRAT"), specialised to
//     b = 1/2 where the incomplete gamma ratio it needs is Q(1/2, z) = erfc(sqrt z).
//     Cost: one log1p, sqrt, exp, erfc and 2..9 recurrence steps.
//   * otherwise: the continued fraction DLMF 8.17.22 (modified Lentz), directly or
//     through I_x(a,b) = 1 - I_{1-x}(b,a).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define NSR_HD __host__ __device__ __forceinline__
#else
#define NSR_HD inline
#endif

#define NSR_BGRAT_TERMS 30

struct NsrPvalParams {
    double a;        // first shape parameter
    double nu;       // a - 1/4
    double ca;       // Gamma(a+1/2) / (Gamma(a) sqrt(nu))
    double v;        // 1 / (4 nu^2)
    double lbeta;    // ln B(a, 1/2)
    int    bgrat;    // a >= 15
};

// d_n of the BGRAT expansion for b = 1/2 (they depend on b only), n = 1..30, from
//   d_n = (b-1) c_n + (1/n) sum_{i<n} (i b - n) c_i d_{n-i},   c_n = 1/(2n+1)!
// evaluated at 60 digits (mpmath) and rounded to float64.
#define NSR_BGRAT_D_TABLE                                                        \
    0.0,                                                                         \
    -8.3333333333333333e-2, 6.25e-3, -5.042989417989418e-4,                      \
    4.343722442680776e-5, -3.8963857323232323e-6, 3.5833354634507908e-7,         \
    -3.3497470720605779e-8, 3.1675854343161148e-9, -3.0210368517370962e-10,      \
    2.900415787669253e-11, -2.7993967483752693e-12, 2.7136393200172037e-13,      \
    -2.6400478921841529e-14, 2.576345276111418e-15, -2.5208132554200556e-16,     \
    2.4721275074517471e-17, -2.4292486395023516e-18, 2.3913480187649295e-19,     \
    -2.3577559464117405e-20, 2.3279246157682319e-21, -2.3014011079062036e-22,    \
    2.2778073573447366e-23, -2.2568250552677534e-24, 2.2381841130778134e-25,     \
    -2.2216537345132669e-26, 2.2070354267468694e-27, -2.1941574717669088e-28,    \
    2.1828705107634888e-29, -2.1730439861920983e-30, 2.1645632514720971e-31

static const double nsr_bgrat_d_host[NSR_BGRAT_TERMS + 1] = {NSR_BGRAT_D_TABLE};
#if defined(__CUDACC__)
static __device__ __constant__ double nsr_bgrat_d_dev[NSR_BGRAT_TERMS + 1] = {NSR_BGRAT_D_TABLE};
#endif

NSR_HD double nsr_bgrat_d(int n) {
#if defined(__CUDA_ARCH__)
    return nsr_bgrat_d_dev[n];
#else
    return nsr_bgrat_d_host[n];
#endif
}

NSR_HD NsrPvalParams nsr_pval_params(double a) {
    NsrPvalParams p;
    p.a = a;
    p.nu = a - 0.25;
    p.v = 0.25 / (p.nu * p.nu);
    p.bgrat = a >= 15.0;
    double lratio;   // ln( Gamma(a+1/2) / Gamma(a) )
    if (a >= 100.0) {
        // sqrt(a) (1 - 1/(8a) + 1/(128a^2) + 5/(1024a^3) - 21/(32768a^4) - 399/(262144a^5))
        double i = 1.0 / a;
        double s = 1.0 + i * (-0.125 + i * (0.0078125 + i * (0.0048828125 +
                   i * (-0.000640869140625 + i * (-0.001522064208984375)))));
        lratio = 0.5 * log(a) + log(s);
    } else {
        lratio = lgamma(a + 0.5) - lgamma(a);
    }
    p.ca = exp(lratio - 0.5 * log(p.nu));
    p.lbeta = 0.57236494292470008707 - lratio;      // ln Gamma(1/2) = ln(pi)/2
    return p;
}

// continued fraction of I_x(a,b) (without the prefactor), modified Lentz
NSR_HD double nsr_betacf(double a, double b, double x) {
    const double tiny = 1e-300;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 20000; ++m) {
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
        if (fabs(del - 1.0) < 2e-16) break;
    }
    return h;
}

// P = I_{1-r2}(a, 1/2)
NSR_HD double nsr_pvalue_r2(double r2_in, const NsrPvalParams& p) {
    // The reference forms x = 1 - R2 in float64 before calling beta.cdf
    // (association.py:249); R2 below 1.1e-16 therefore gives exactly P = 1.  Keep
    // that rounding so tiny correlations agree with it to the last digit.
    const double x = 1.0 - r2_in;
    if (x != x) return x;                          // NaN in, NaN out (the reference asserts on it)
    if (!(x < 1.0)) return 1.0;
    if (!(x > 0.0)) return 0.0;
    const double r2 = 1.0 - x;                     // exact (Sterbenz) for x >= 1/2
    const double lnx = log1p(-r2);                 // ln x
    if (p.bgrat && r2 < 0.3) {
        // I_x(a,1/2) ~ ca [ Q(1/2,z) + R sum_n d_n J_n ],  z = -nu ln x,  Q = erfc(sqrt z),
        // R = sqrt(z) e^-z / sqrt(pi), J_0 = Q / R.  With erfcx(s) = e^(s^2) erfc(s) everything
        // is e^-z times well-scaled factors: no 0/0 when e^-z underflows.
        const double z = -p.nu * lnx;
        const double sz = sqrt(z);
        const double ez = exp(-z);
        const double ex = erfcx(sz);
        const double t2 = 0.25 * lnx * lnx;
        double j = ex * (1.7724538509055160273 / sz), t = 1.0, n2 = 0.0, s = 0.0;
        const double j0 = j;
#pragma unroll 1
        for (int n = 1; n <= NSR_BGRAT_TERMS; ++n) {
            const double bp2n = 0.5 + n2;
            j = (bp2n * (bp2n + 1.0) * j + (z + bp2n + 1.0) * t) * p.v;
            n2 += 2.0;
            t *= t2;
            const double dj = nsr_bgrat_d(n) * j;
            s += dj;
            if (fabs(dj) <= 1e-17 * (j0 + s)) break;
        }
        return fmin(1.0, p.ca * ez * (ex + sz * 0.56418958354775628695 * s));
    }
    // general route
    const double lpre = p.a * lnx + 0.5 * log(r2) - p.lbeta;      // ln[x^a (1-x)^b / B(a,b)]
    if (x < (p.a + 1.0) / (p.a + 2.5))
        return fmin(1.0, exp(lpre) * nsr_betacf(p.a, 0.5, x) / p.a);
    return fmax(0.0, 1.0 - exp(lpre) * nsr_betacf(0.5, p.a, r2) * 2.0);
}

// Two P-values in lockstep: the float64 dependency chains of the expansion are latency bound, so
// evaluating two independent elements together nearly doubles the epilogue's throughput.
// Results are bit-identical to two nsr_pvalue_r2 calls.
NSR_HD void nsr_pvalue_r2_x2(double r2a_in, double r2b_in, const NsrPvalParams& p, double& pa, double& pb) {
    const double xa = 1.0 - r2a_in, xb = 1.0 - r2b_in;
    const double r2a = 1.0 - xa, r2b = 1.0 - xb;
    const bool oka = p.bgrat && xa < 1.0 && xa > 0.0 && r2a < 0.3;
    const bool okb = p.bgrat && xb < 1.0 && xb > 0.0 && r2b < 0.3;
    if (!(oka && okb)) {
        pa = nsr_pvalue_r2(r2a_in, p);
        pb = nsr_pvalue_r2(r2b_in, p);
        return;
    }
    const double lnxa = log1p(-r2a), lnxb = log1p(-r2b);
    const double za = -p.nu * lnxa, zb = -p.nu * lnxb;
    const double sza = sqrt(za), szb = sqrt(zb);
    const double eza = exp(-za), ezb = exp(-zb);
    const double exa = erfcx(sza), exb = erfcx(szb);
    const double t2a = 0.25 * lnxa * lnxa, t2b = 0.25 * lnxb * lnxb;
    double ja = exa * (1.7724538509055160273 / sza), jb = exb * (1.7724538509055160273 / szb);
    double ta = 1.0, tb = 1.0, n2 = 0.0, sa = 0.0, sb = 0.0;
    const double j0a = ja, j0b = jb;
    bool donea = false, doneb = false;
#pragma unroll 1
    for (int n = 1; n <= NSR_BGRAT_TERMS; ++n) {
        const double bp2n = 0.5 + n2;
        const double c1 = bp2n * (bp2n + 1.0);
        const double d = nsr_bgrat_d(n);
        ja = (c1 * ja + (za + bp2n + 1.0) * ta) * p.v;
        jb = (c1 * jb + (zb + bp2n + 1.0) * tb) * p.v;
        n2 += 2.0;
        ta *= t2a;
        tb *= t2b;
        const double dja = d * ja, djb = d * jb;
        if (!donea) { sa += dja; donea = fabs(dja) <= 1e-17 * (j0a + sa); }
        if (!doneb) { sb += djb; doneb = fabs(djb) <= 1e-17 * (j0b + sb); }
        if (donea && doneb) break;
    }
    pa = fmin(1.0, p.ca * eza * (exa + sza * 0.56418958354775628695 * sa));
    pb = fmin(1.0, p.ca * ezb * (exb + szb * 0.56418958354775628695 * sb));
}
