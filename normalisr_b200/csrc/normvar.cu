// Variance normalisation, the step directly upstream of coex / de.
// Reference: src/normalisr/norm.py:131-289 (normvar / normvar1): gene x is scaled by
// s_k = w_k ** wt_x per cell and the covariates dc * s (the gene's OWN weighted covariates) are
// projected out of it, one pseudo-inverse per gene; the reference loops over genes in Python.
//
// Two streaming passes over dt:
//   nsr_normvar_stats   per gene: G = sum_k s^2 dc dc^T (upper triangle), b = sum_k s^2 dc dt,
//                       S1 = sum_k s dt, S2 = sum_k (s dt)^2        (float64 FMA pipe)
//   (host layer: pseudo-inverse of every nc x nc G with the reference's rank rule, coef = G+ b,
//    residual variance S2 - b^T G+ b, keepvar scale)
//   nsr_normvar_apply   out = scale * s * (dt - coef^T dc)
// A warp owns a gene and strides over the cells (coalesced 8-byte reads of dt); the covariate
// chunk and log w are staged in shared memory once per CTA of 8 genes.  s = exp(wt * log w).
#include "nsr_common.cuh"

namespace {

constexpr int kNvThreads = 256;
constexpr int kNvWarps = kNvThreads / 32;
constexpr int kNvChunk = 384;          // cells staged per step (12 x 384 doubles + log w < 48 KB static)
constexpr int kNvAhead = 4;            // cells per lane whose loads are issued before any arithmetic

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

template <int NC>
__global__ void __launch_bounds__(kNvThreads)
normvar_stats_kernel(const double* __restrict__ dt, int64_t genes, int64_t n, int64_t ld,
                     const double* __restrict__ dc, int nc, int64_t ldc, const double* __restrict__ logw,
                     const double* __restrict__ wt, double* __restrict__ stats) {
    __shared__ double s_c[NC][kNvChunk];
    __shared__ double s_lw[kNvChunk];
    constexpr int kTri = NC * (NC + 1) / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gene = (int64_t)blockIdx.x * kNvWarps + warp;
    const bool live = gene < genes;
    const double wtx = live ? wt[gene] : 0.0;
    const double* row = dt + (live ? gene : 0) * ld;
    double g[kTri], b[NC], s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int i = 0; i < kTri; ++i) g[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) b[i] = 0.0;

    for (int64_t k0 = 0; k0 < n; k0 += kNvChunk) {
        const int len = (int)min((int64_t)kNvChunk, n - k0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < NC * kNvChunk; idx += kNvThreads) {
            const int j = idx / kNvChunk, k = idx % kNvChunk;
            s_c[j][k] = (j < nc && k < len) ? dc[(int64_t)j * ldc + k0 + k] : 0.0;
        }
        for (int k = threadIdx.x; k < kNvChunk; k += kNvThreads) s_lw[k] = k < len ? logw[k0 + k] : 0.0;
        __syncthreads();
        if (live) {
            for (int kq = lane; kq < len; kq += 32 * kNvAhead) {
                // issue the loads of kNvAhead cells first: 8 resident warps per SM cannot hide HBM latency otherwise
                double x[kNvAhead], lw[kNvAhead];
#pragma unroll
                for (int u = 0; u < kNvAhead; ++u) {
                    const int k = kq + 32 * u;
                    x[u] = k < len ? row[k0 + k] : 0.0;
                    lw[u] = s_lw[k < kNvChunk ? k : 0];
                }
#pragma unroll
                for (int u = 0; u < kNvAhead; ++u) {
                    const int k = kq + 32 * u;
                    if (k < len) {
                        const double s = wtx == 0.0 ? 1.0 : exp(wtx * lw[u]);  // norm.py:238-239
                        const double v = x[u] * s;
                        s1 += v;
                        s2 = fma(v, v, s2);
                        double a[NC];
#pragma unroll
                        for (int j = 0; j < NC; ++j) a[j] = s_c[j][k] * s;
                        int t = 0;
#pragma unroll
                        for (int i = 0; i < NC; ++i) {
                            b[i] = fma(a[i], v, b[i]);
#pragma unroll
                            for (int j = i; j < NC; ++j) { g[t] = fma(a[i], a[j], g[t]); ++t; }
                        }
                    }
                }
            }
        }
    }
    if (!live) return;
    // fixed-shape butterfly over the lanes
    double* o = stats + gene * (kTri + NC + 2);
#pragma unroll
    for (int i = 0; i < kTri; ++i) {
        const double v = warp_sum(g[i]);
        if (lane == 0) o[i] = v;
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const double v = warp_sum(b[i]);
        if (lane == 0) o[kTri + i] = v;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { o[kTri + NC] = s1; o[kTri + NC + 1] = s2; }
}

template <int NC>
__global__ void __launch_bounds__(kNvThreads)
normvar_apply_kernel(const double* __restrict__ dt, int64_t genes, int64_t n, int64_t ld,
                     const double* __restrict__ dc, int nc, int64_t ldc, const double* __restrict__ logw,
                     const double* __restrict__ wt, const double* __restrict__ coef, const double* __restrict__ scale,
                     double* __restrict__ out, int64_t ldo) {
    __shared__ double s_c[NC][kNvChunk];
    __shared__ double s_lw[kNvChunk];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gene = (int64_t)blockIdx.x * kNvWarps + warp;
    const bool live = gene < genes;
    const double wtx = live ? wt[gene] : 0.0;
    const double sc = live ? scale[gene] : 0.0;
    double c[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) c[j] = (live && j < nc) ? coef[gene * nc + j] : 0.0;
    const double* row = dt + (live ? gene : 0) * ld;
    double* orow = out + (live ? gene : 0) * ldo;
    // cells are split over blockIdx.y so that the grid fills the machine for any gene count
    const int64_t per = ((n + gridDim.y - 1) / gridDim.y + kNvChunk - 1) / kNvChunk * kNvChunk;
    const int64_t kb = (int64_t)blockIdx.y * per, ke = min(n, kb + per);
    for (int64_t k0 = kb; k0 < ke; k0 += kNvChunk) {
        const int len = (int)min((int64_t)kNvChunk, ke - k0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < NC * kNvChunk; idx += kNvThreads) {
            const int j = idx / kNvChunk, k = idx % kNvChunk;
            s_c[j][k] = (j < nc && k < len) ? dc[(int64_t)j * ldc + k0 + k] : 0.0;
        }
        for (int k = threadIdx.x; k < kNvChunk; k += kNvThreads) s_lw[k] = k < len ? logw[k0 + k] : 0.0;
        __syncthreads();
        if (live) {
            for (int kq = lane; kq < len; kq += 32 * kNvAhead) {
                double x[kNvAhead];
#pragma unroll
                for (int u = 0; u < kNvAhead; ++u) {
                    const int k = kq + 32 * u;
                    x[u] = k < len ? row[k0 + k] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kNvAhead; ++u) {
                    const int k = kq + 32 * u;
                    if (k < len) {
                        const double s = wtx == 0.0 ? 1.0 : exp(wtx * s_lw[k]);
                        double r = x[u];
#pragma unroll
                        for (int j = 0; j < NC; ++j) r = fma(-c[j], s_c[j][k], r);
                        orow[k0 + k] = sc * (s * r);
                    }
                }
            }
        }
    }
}

}  // namespace

#define NSR_NV_DISPATCH(NCV, CALL)                                  \
    do {                                                            \
        if ((NCV) <= 4) { CALL(4); }                                \
        else if ((NCV) <= 6) { CALL(6); }                           \
        else if ((NCV) <= 8) { CALL(8); }                           \
        else if ((NCV) <= 10) { CALL(10); }                         \
        else { CALL(12); }                                          \
    } while (0)

extern "C" int nsr_normvar_width(int nc) {
    if (nc < 1 || nc > 12) return -1;
    return nc <= 4 ? 4 : nc <= 6 ? 6 : nc <= 8 ? 8 : nc <= 10 ? 10 : 12;
}

extern "C" int nsr_normvar_stats(nsr_ctx* ctx, uintptr_t stream, const double* dt, int64_t genes, int64_t n,
                                 int64_t ld, const double* dc, int nc, int64_t ldc, const double* logw,
                                 const double* wt, double* stats) {
    NSR_REQUIRE(ctx && dt && dc && logw && wt && stats, "nsr_normvar_stats: null argument");
    NSR_REQUIRE(genes >= 1 && n >= 1 && ld >= n && ldc >= n && nc >= 1 && nc <= 12,
                "nsr_normvar_stats: bad shape genes=%lld n=%lld nc=%d (1..12 covariates)", (long long)genes,
                (long long)n, nc);
    NSR_CHECK(cudaSetDevice(ctx->device));
    const unsigned grid = (unsigned)((genes + kNvWarps - 1) / kNvWarps);
    cudaStream_t st = (cudaStream_t)stream;
#define NSR_NV_STATS(W) normvar_stats_kernel<W><<<grid, kNvThreads, 0, st>>>(dt, genes, n, ld, dc, nc, ldc, logw, wt, stats)
    NSR_NV_DISPATCH(nc, NSR_NV_STATS);
#undef NSR_NV_STATS
    NSR_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int nsr_normvar_apply(nsr_ctx* ctx, uintptr_t stream, const double* dt, int64_t genes, int64_t n,
                                 int64_t ld, const double* dc, int nc, int64_t ldc, const double* logw,
                                 const double* wt, const double* coef, const double* scale, double* out,
                                 int64_t ldo) {
    NSR_REQUIRE(ctx && dt && dc && logw && wt && coef && scale && out, "nsr_normvar_apply: null argument");
    NSR_REQUIRE(genes >= 1 && n >= 1 && ld >= n && ldo >= n && ldc >= n && nc >= 1 && nc <= 12,
                "nsr_normvar_apply: bad shape genes=%lld n=%lld nc=%d (1..12 covariates)", (long long)genes,
                (long long)n, nc);
    NSR_CHECK(cudaSetDevice(ctx->device));
    const int64_t gx = (genes + kNvWarps - 1) / kNvWarps;
    int64_t gy = (4 * (int64_t)ctx->sm_count + gx - 1) / gx;                 // >= 4 CTAs per SM in total
    const int64_t max_y = (n + kNvChunk - 1) / kNvChunk;
    if (gy > max_y) gy = max_y;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const dim3 grid((unsigned)gx, (unsigned)gy);
    cudaStream_t st = (cudaStream_t)stream;
#define NSR_NV_APPLY(W) normvar_apply_kernel<W><<<grid, kNvThreads, 0, st>>>(dt, genes, n, ld, dc, nc, ldc, logw, wt, coef, scale, out, ldo)
    NSR_NV_DISPATCH(nc, NSR_NV_APPLY);
#undef NSR_NV_APPLY
    NSR_CHECK(cudaGetLastError());
    return 0;
}
