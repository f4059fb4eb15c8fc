// Host build of csrc/pvalue.cuh for tests/test_pvalue_host.py (no GPU needed).
// glibc has no erfcx(); CUDA does.  Stand-in, used by this test build only.
#include <math.h>
#include <stdint.h>
static inline double erfcx(double x) {
    if (x < 25.0) return exp(x * x) * erfc(x);
    const double i2 = 1.0 / (2.0 * x * x);
    return (1.0 / (x * 1.7724538509055160273)) *
           (1.0 - i2 * (1.0 - 3.0 * i2 * (1.0 - 5.0 * i2 * (1.0 - 7.0 * i2 * (1.0 - 9.0 * i2)))));
}
#include "pvalue.cuh"
extern "C" void pv_each(const double* r2, const double* a, double* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        NsrPvalParams p = nsr_pval_params(a[i]);
        out[i] = nsr_pvalue_r2(r2[i], p);
    }
}
extern "C" void pv_pairs(const double* r2, double a, double* out, int64_t n) {
    NsrPvalParams p = nsr_pval_params(a);
    for (int64_t i = 0; i + 1 < n; i += 2) nsr_pvalue_r2_x2(r2[i], r2[i + 1], p, out[i], out[i + 1]);
}
