// Variance normalisation, the step directly upstream of coex / de.
// Reference: src/normalisr/norm.py:131-289 (normvar / normvar1): gene x is scaled by
// s_k = w_k ** wt_x per cell and the covariates dc * s (the gene's OWN weighted covariates) are
// projected out of it, one pseudo-inverse per gene; the reference loops over genes in Python.
//
// Two streaming passes over dt:
//   nsr_normvar_stats   per gene: G = sum_k s^2 dc dc^T (upper triangle), b = sum_k s^2 dc dt,
//                       S1 = sum_k s dt, S2 = sum_k (s dt)^2        (FP64 tensor cores, see below)
//   (host layer: pseudo-inverse of every nc x nc G with the reference's rank rule, coef = G+ b,
//    residual variance S2 - b^T G+ b, keepvar scale)
//   nsr_normvar_apply   out = scale * s * (dt - coef^T dc)
// s = exp(wt * log w).  In the apply pass a warp owns a gene and strides over the cells (coalesced
// 8-byte accesses); the covariate chunk and log w are staged in shared memory once per CTA of 8 genes.
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

// ---- pass 1 on the FP64 tensor cores -------------------------------------------------------
// G_x[i][j] = sum_k s_k^2 (c_ik c_jk) and b_x[i] = sum_k (s_k^2 dt_xk) c_ik are products of a
// (genes x cells) matrix that depends on the gene only through s = w ** wt_x with gene-INDEPENDENT
// (columns x cells) matrices: D (the nc (nc + 1) / 2 products c_i c_j, built once by the host layer)
// and C.  So the statistics are two skinny GEMMs over cells, M = genes, K = cells:
//     [G | b] = [ s^2 | s^2 dt ] x [ D | C ]^T          (mma.sync m8n8k4 f64, "DMMA")
// with the A operands generated on the fly (one exp per matrix entry).  A warp owns 16 genes
// (2 row tiles) x one cell split; fragment layout as in residual.cu (g = lane / 4, t = lane % 4:
// lane (g, t) feeds cells 16 kk + 4 t + e of row g to MMA e).  S1 = sum s dt and S2 = sum (s dt)^2
// are plain per-lane sums.  Partial sums over the n-only cell splits are combined in a fixed order.
constexpr int kGmWarps = 8;
constexpr int kGmTiles = 2;            // row tiles (of 8 genes) per warp

__device__ __forceinline__ void nv_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void nv_load4(const double* p, int64_t k, int64_t n, double (&v)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (k + e < n) ? __ldg(p + e) : 0.0;
}

// M: (8 * (NTD + NTC) x n) = [D rows, zero rows up to 8 NTD | C rows, zero rows up to 8 NTC]
template <int NTD, int NTC>
__global__ void __launch_bounds__(32 * kGmWarps)
normvar_gemm_kernel(const double* __restrict__ dt, int64_t genes, int64_t n, int64_t ld,
                    const double* __restrict__ M, int64_t ldm, const double* __restrict__ logw,
                    const double* __restrict__ wt, int ksplit, double* __restrict__ partial) {
    constexpr int NT = NTD + NTC;
    constexpr int kCols = 8 * NT + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t gene_w = ((int64_t)blockIdx.x * kGmWarps + warp) * (8 * kGmTiles);
    if (gene_w >= genes) return;
    const double* xr[kGmTiles];
    double wtx[kGmTiles];
    bool valid[kGmTiles];
#pragma unroll
    for (int r = 0; r < kGmTiles; ++r) {
        const int64_t gene = gene_w + 8 * r + g;
        valid[r] = gene < genes;
        xr[r] = dt + (valid[r] ? gene : gene_w) * ld + 4 * t;
        wtx[r] = valid[r] ? wt[gene] : 0.0;
    }
    const int64_t n16 = (n + 15) / 16;
    const int64_t kb = n16 * blockIdx.y / ksplit, ke = n16 * (blockIdx.y + 1) / ksplit;
    double acc[kGmTiles][NT][2], s1[kGmTiles], s2[kGmTiles];
#pragma unroll
    for (int r = 0; r < kGmTiles; ++r) {
        s1[r] = s2[r] = 0.0;
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[r][j][0] = acc[r][j][1] = 0.0;
    }
#pragma unroll 1
    for (int64_t kk = kb; kk < ke; ++kk) {
        const int64_t k = kk * 16 + 4 * t;
        double xv[kGmTiles][4], lw[4], mv[NT][4];
#pragma unroll
        for (int r = 0; r < kGmTiles; ++r) nv_load4(xr[r] + kk * 16, k, valid[r] ? n : 0, xv[r]);
        nv_load4(logw + k, k, n, lw);
#pragma unroll
        for (int j = 0; j < NT; ++j) nv_load4(M + (int64_t)(8 * j + g) * ldm + k, k, n, mv[j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int r = 0; r < kGmTiles; ++r) {
                const double s = wtx[r] == 0.0 ? 1.0 : exp(wtx[r] * lw[e]);     // norm.py:238-239
                const double v = xv[r][e] * s;
                const double a1 = s * s, a2 = s * v;
                s1[r] += v;
                s2[r] = fma(v, v, s2[r]);
#pragma unroll
                for (int j = 0; j < NTD; ++j) nv_dmma(acc[r][j][0], acc[r][j][1], a1, mv[j][e]);
#pragma unroll
                for (int j = NTD; j < NT; ++j) nv_dmma(acc[r][j][0], acc[r][j][1], a2, mv[j][e]);
            }
        }
    }
    // accumulator layout: row g, columns 8 j + 2 t, + 1
#pragma unroll
    for (int r = 0; r < kGmTiles; ++r) {
        const int64_t gene = gene_w + 8 * r + g;
        double* o = partial + ((int64_t)blockIdx.y * genes + (valid[r] ? gene : 0)) * kCols;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            if (valid[r]) { o[8 * j + 2 * t] = acc[r][j][0]; o[8 * j + 2 * t + 1] = acc[r][j][1]; }
        }
        double a = s1[r], b = s2[r];
        a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
        b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
        if (valid[r] && t == 0) { o[8 * NT] = a; o[8 * NT + 1] = b; }
    }
}

// stats[gene][c] = sum over cell splits, fixed order
__global__ void normvar_reduce_kernel(const double* __restrict__ partial, int64_t genes, int cols, int ksplit,
                                      double* __restrict__ stats) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= genes * cols) return;
    double s = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) s += partial[(int64_t)ks * genes * cols + i];
    stats[i] = s;
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

// padded column count of the statistics for nc covariates: 8 * (D tiles + C tiles) + 2
static int nv_d_tiles(int nc) {
    const int need = (nc * (nc + 1) / 2 + 7) / 8;
    return need <= 2 ? 2 : need <= 4 ? 4 : need <= 6 ? 6 : need <= 8 ? 8 : 10;
}
extern "C" int nsr_normvar_width(int nc) {
    if (nc < 1 || nc > 12) return -1;
    return 8 * (nv_d_tiles(nc) + 2) + 2;
}

extern "C" int nsr_normvar_stats(nsr_ctx* ctx, uintptr_t stream, const double* dt, int64_t genes, int64_t n,
                                 int64_t ld, const double* M, int nc, int64_t ldm, const double* logw,
                                 const double* wt, double* stats) {
    NSR_REQUIRE(ctx && dt && M && logw && wt && stats, "nsr_normvar_stats: null argument");
    NSR_REQUIRE(genes >= 1 && n >= 1 && ld >= n && ldm >= n && nc >= 1 && nc <= 12,
                "nsr_normvar_stats: bad shape genes=%lld n=%lld nc=%d (1..12 covariates)", (long long)genes,
                (long long)n, nc);
    NSR_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int ntd = nv_d_tiles(nc), cols = 8 * (ntd + 2) + 2;
    const int ksplit = nsr_cell_splits(n);
    void* scratch = nullptr;
    if (nsr_scratch(ctx, (size_t)ksplit * genes * cols * sizeof(double), &scratch)) return 1;
    const dim3 grid((unsigned)((genes + kGmWarps * 8 * kGmTiles - 1) / (kGmWarps * 8 * kGmTiles)), (unsigned)ksplit);
#define NSR_NV_GEMM(D_) normvar_gemm_kernel<D_, 2><<<grid, 32 * kGmWarps, 0, st>>>(dt, genes, n, ld, M, ldm, logw, wt, ksplit, (double*)scratch)
    switch (ntd) {
        case 2: NSR_NV_GEMM(2); break;
        case 4: NSR_NV_GEMM(4); break;
        case 6: NSR_NV_GEMM(6); break;
        case 8: NSR_NV_GEMM(8); break;
        default: NSR_NV_GEMM(10); break;
    }
#undef NSR_NV_GEMM
    const int64_t total = genes * cols;
    normvar_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const double*)scratch, genes, cols, ksplit, stats);
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
