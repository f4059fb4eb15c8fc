// Shared epilogue of the contraction kernels: exact integer cell-sums -> dot, r^2, P.
// Reference: association.py:234-249 (gamma, R2, beta.cdf) and :1036-1057 (dot, symmetry).
#pragma once
#include "nsr_common.cuh"
#include "pvalue.cuh"

struct ContractParams {
    int mode;                    // NSR_MODE_*
    int n_groups;                // weight groups kept (w = 2 .. n_groups + 1)
    int64_t rows_a, rows_b, ld;
    const double* qa; const double* va;
    const double* qb; const double* vb;
    double* P; double* out2;
    double inv_n;
    double group_scale[4];       // 256^(2S - w), relative to the least significant kept group
    double scale_all;            // weight of the least significant kept group
    NsrPvalParams pv;
};

// acc = exact sum_k V_ik V_jk restricted to the kept digit products, as a double
__device__ __forceinline__ double nsr_combine(const ContractParams& p, const int32_t* acc_g) {
    double s = 0.0;
#pragma unroll
    for (int g = 0; g < 4; ++g)
        if (g < p.n_groups) s = fma((double)acc_g[g], p.group_scale[g], s);
    return s * p.scale_all;
}

// one output element (i = row in A, j = row in B); `mirror` also writes (j, i) (COEX, i != j tile)
__device__ __forceinline__ void nsr_finish(const ContractParams& p, int64_t i, int64_t j, double qi,
                                           double vi, double qj, double vj, double acc,
                                           bool mirror) {
    const double sum = (qi * qj) * acc;            // sum_k res_i res_j
    if (p.mode == NSR_MODE_RAW) {
        p.out2[i * p.ld + j] = sum;
        return;
    }
    const double dot = sum * p.inv_n;
    double P, o2;
    if ((p.mode == NSR_MODE_COEX || p.mode == NSR_MODE_COEX_UPPER) && i == j) {
        P = 0.0; o2 = 0.0;                         // triu(.,1) + transpose leaves a zero diagonal
    } else {
        const double r2 = (dot * dot) / (vi * vj);
        P = nsr_pvalue_r2(r2, p.pv);
        o2 = (p.mode == NSR_MODE_DE) ? dot / vi : dot;
    }
    p.P[i * p.ld + j] = P;
    p.out2[i * p.ld + j] = o2;
    if (mirror) {
        p.P[j * p.ld + i] = P;
        p.out2[j * p.ld + i] = o2;
    }
}

// Out-of-line pair evaluation: keeps the fully unrolled 64-element epilogue loops small.
static __device__ __noinline__ double2 nsr_pvalue_pair(double r2a, double r2b, const NsrPvalParams* p) {
    double2 r;
    nsr_pvalue_r2_x2(r2a, r2b, *p, r.x, r.y);
    return r;
}

// Two adjacent output elements (i, j) and (i, j+1); j+1 may be out of range (second = false).
__device__ __forceinline__ void nsr_finish2(const ContractParams& p, int64_t i, int64_t j, double qi, double vi,
                                            double acc0, double acc1, bool second, bool mirror) {
    const double qj0 = p.qb[j], qj1 = second ? p.qb[j + 1] : 0.0;
    const double sum0 = (qi * qj0) * acc0, sum1 = (qi * qj1) * acc1;
    double* o2 = p.out2 + i * p.ld + j;
    if (p.mode == NSR_MODE_RAW) {
        o2[0] = sum0;
        if (second) o2[1] = sum1;
        return;
    }
    const double vj0 = p.vb[j], vj1 = second ? p.vb[j + 1] : 1.0;
    const double dot0 = sum0 * p.inv_n, dot1 = sum1 * p.inv_n;
    const bool sym = p.mode == NSR_MODE_COEX || p.mode == NSR_MODE_COEX_UPPER;
    double2 P = nsr_pvalue_pair((dot0 * dot0) / (vi * vj0), (dot1 * dot1) / (vi * vj1), &p.pv);
    double w0 = (p.mode == NSR_MODE_DE) ? dot0 / vi : dot0;
    double w1 = (p.mode == NSR_MODE_DE) ? dot1 / vi : dot1;
    if (sym && i == j) { P.x = 0.0; w0 = 0.0; }        // zero diagonal (association.py:1049-1057)
    if (sym && i == j + 1) { P.y = 0.0; w1 = 0.0; }
    double* pp = p.P + i * p.ld + j;
    pp[0] = P.x;
    o2[0] = w0;
    if (second) { pp[1] = P.y; o2[1] = w1; }
    if (mirror) {
        p.P[j * p.ld + i] = P.x;
        p.out2[j * p.ld + i] = w0;
        if (second) {
            p.P[(j + 1) * p.ld + i] = P.y;
            p.out2[(j + 1) * p.ld + i] = w1;
        }
    }
}
