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
// s = w ** wt = exp(wt * log w).  Both passes run on the FP64 tensor cores (DMMA) in the same fragment
// layout; see the kernels.
#include "nsr_common.cuh"

namespace {

constexpr int kNvThreads = 256;
constexpr int kNvWarps = kNvThreads / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// s = w ** wt = 2^(wt log2(e) log w) without the library's special-case handling: round to the nearest
// integer k, e^z on |z| <= ln(2)/2 by its Taylor polynomial of degree 13 (truncation 4e-18), 2^k added to
// the exponent field.  0.7 ulp against exp2l on [-30, 30] (host twin of this code); arguments beyond
// +-1000 (over / underflow territory) take the library path.  wt2 = wt * log2(e), once per gene.
constexpr double kLog2e = 1.4426950408889634;
__device__ __forceinline__ double nv_exp2(double y) {
    if (!(fabs(y) < 1000.0)) return exp2(y);
    const double kf = rint(y);
    const double z = (y - kf) * 0.6931471805599453;
    double p = 1.6059043836821613e-10;            // 1/13!
    p = fma(p, z, 2.08767569878681e-09);
    p = fma(p, z, 2.505210838544172e-08);
    p = fma(p, z, 2.755731922398589e-07);
    p = fma(p, z, 2.7557319223985893e-06);
    p = fma(p, z, 2.48015873015873e-05);
    p = fma(p, z, 1.984126984126984e-04);
    p = fma(p, z, 1.388888888888889e-03);
    p = fma(p, z, 8.333333333333333e-03);
    p = fma(p, z, 4.1666666666666664e-02);
    p = fma(p, z, 1.6666666666666666e-01);
    p = fma(p, z, 0.5);
    p = fma(p, z, 1.0);
    p = fma(p, z, 1.0);
    return __hiloint2double(__double2hiint(p) + ((int)kf << 20), __double2loint(p));
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
// the same four values with two 16-byte loads (p 16-byte aligned, all four inside the row)
__device__ __forceinline__ void nv_load4v(const double* p, double (&v)[4]) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// M: (8 * (NTD + NTC) x n) = [D rows, zero rows up to 8 NTD | C rows, zero rows up to 8 NTC]
template <int NTD, int NTC, bool VEC = false>
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
        wtx[r] = valid[r] ? wt[gene] * kLog2e : 0.0;
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
        if (VEC && kk * 16 + 16 <= n) {              // interior step, 16-byte aligned rows: no predicates
#pragma unroll
            for (int r = 0; r < kGmTiles; ++r) {
                if (valid[r]) nv_load4v(xr[r] + kk * 16, xv[r]);
                else xv[r][0] = xv[r][1] = xv[r][2] = xv[r][3] = 0.0;
            }
            nv_load4v(logw + k, lw);
#pragma unroll
            for (int j = 0; j < NT; ++j) nv_load4v(M + (int64_t)(8 * j + g) * ldm + k, mv[j]);
        } else {
#pragma unroll
            for (int r = 0; r < kGmTiles; ++r) nv_load4(xr[r] + kk * 16, k, valid[r] ? n : 0, xv[r]);
            nv_load4(logw + k, k, n, lw);
#pragma unroll
            for (int j = 0; j < NT; ++j) nv_load4(M + (int64_t)(8 * j + g) * ldm + k, k, n, mv[j]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int r = 0; r < kGmTiles; ++r) {
                const double s = wtx[r] == 0.0 ? 1.0 : nv_exp2(wtx[r] * lw[e]);     // w ** wt, norm.py:238-239
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

// pass 2 on the FP64 tensor cores: out = scale * s * (dt - coef^T dc).  The covariate part is a
// (genes x covariates) x (covariates x cells) product: A = -coef (a warp's 16 genes, constant over its
// cell loop, in registers), B = dc (4 covariates x 8 cells per DMMA, read through L1: dc is a few MB and
// shared by every warp), accumulator initialised with the dt tile - so the residual comes out of the MMA
// in the accumulator layout (lane (g, t): gene g, cells 2 t and 2 t + 1), where it is multiplied by
// s = w ** wt and the keepvar scale and stored.  KS = ceil(nc / 4) k-steps.  No shared memory, no barrier.
template <int KS, bool VEC>
__global__ void __launch_bounds__(kNvThreads)
normvar_apply_kernel(const double* __restrict__ dt, int64_t genes, int64_t n, int64_t ld,
                     const double* __restrict__ dc, int nc, int64_t ldc, const double* __restrict__ logw,
                     const double* __restrict__ wt, const double* __restrict__ coef, const double* __restrict__ scale,
                     double* __restrict__ out, int64_t ldo, int64_t cells_per_split) {
    constexpr int R = 2, U = 2;                   // row tiles (8 genes) and cell tiles (8 cells) per step
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t gene_w = ((int64_t)blockIdx.x * kNvWarps + warp) * (8 * R);
    if (gene_w >= genes) return;
    double a[R][KS], wt2[R], sc[R];
    const double* xr[R];
    double* orow[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t gene = gene_w + 8 * r + g;
        valid[r] = gene < genes;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            const int j = 4 * kk + t;
            a[r][kk] = (valid[r] && j < nc) ? -coef[gene * nc + j] : 0.0;
        }
        wt2[r] = valid[r] ? wt[gene] * kLog2e : 0.0;
        sc[r] = valid[r] ? scale[gene] : 0.0;
        xr[r] = dt + (valid[r] ? gene : gene_w) * ld + 2 * t;
        orow[r] = out + (valid[r] ? gene : gene_w) * ldo + 2 * t;
    }
    const bool rows_ok = gene_w + 8 * R <= genes;        // warp-uniform: the MMAs below need all 32 lanes on one path
    const int64_t kb = (int64_t)blockIdx.y * cells_per_split, ke = min(n, kb + cells_per_split);
#pragma unroll 1
    for (int64_t c0 = kb; c0 < ke; c0 += 8 * U) {
        double bf[U][KS], lw[U][2], d[R][U][2];
        if (VEC && rows_ok && c0 + 8 * U <= ke) {
            // interior step: every row and cell exists - 16-byte loads and stores, no predicates
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    const int j = 4 * kk + t;
                    bf[u][kk] = j < nc ? __ldg(dc + (int64_t)j * ldc + c0 + 8 * u + g) : 0.0;
                }
                const double2 l2 = __ldg(reinterpret_cast<const double2*>(logw + c0 + 8 * u + 2 * t));
                lw[u][0] = l2.x; lw[u][1] = l2.y;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double2 v = *reinterpret_cast<const double2*>(xr[r] + c0 + 8 * u);
                    d[r][u][0] = v.x; d[r][u][1] = v.y;
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int kk = 0; kk < KS; ++kk) nv_dmma(d[r][u][0], d[r][u][1], a[r][kk], bf[u][kk]);
                    double2 o;
                    o.x = sc[r] * ((wt2[r] == 0.0 ? 1.0 : nv_exp2(wt2[r] * lw[u][0])) * d[r][u][0]);
                    o.y = sc[r] * ((wt2[r] == 0.0 ? 1.0 : nv_exp2(wt2[r] * lw[u][1])) * d[r][u][1]);
                    *reinterpret_cast<double2*>(orow[r] + c0 + 8 * u) = o;
                }
            }
            continue;
        }
        bool in[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t cb = c0 + 8 * u + g;                    // B operand: column (cell) g of the tile
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const int j = 4 * kk + t;
                bf[u][kk] = (j < nc && cb < ke) ? __ldg(dc + (int64_t)j * ldc + cb) : 0.0;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t k = c0 + 8 * u + 2 * t + e;         // accumulator: cells 2 t, 2 t + 1
                in[u][e] = k < ke;
                lw[u][e] = in[u][e] ? __ldg(logw + k) : 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) d[r][u][e] = (in[u][e] && valid[r]) ? xr[r][c0 + 8 * u + e] : 0.0;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) nv_dmma(d[r][u][0], d[r][u][1], a[r][kk], bf[u][kk]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double s = wt2[r] == 0.0 ? 1.0 : nv_exp2(wt2[r] * lw[u][e]);
                    if (in[u][e] && valid[r]) orow[r][c0 + 8 * u + e] = sc[r] * (s * d[r][u][e]);
                }
            }
        }
    }
}

}  // namespace

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

extern "C" int nsr_normvar_rhs(nsr_ctx* ctx, uintptr_t stream, const double* dt, int64_t genes, int64_t n,
                               int64_t ld, const double* C16, int64_t ldc, const double* logw, const double* wt,
                               double* stats) {
    NSR_REQUIRE(ctx && dt && C16 && logw && wt && stats, "nsr_normvar_rhs: null argument");
    NSR_REQUIRE(genes >= 1 && n >= 1 && ld >= n && ldc >= n, "nsr_normvar_rhs: bad shape genes=%lld n=%lld",
                (long long)genes, (long long)n);
    NSR_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int cols = 18;
    const int ksplit = nsr_cell_splits(n);
    void* scratch = nullptr;
    if (nsr_scratch(ctx, (size_t)ksplit * genes * cols * sizeof(double), &scratch)) return 1;
    const dim3 grid((unsigned)((genes + kGmWarps * 8 * kGmTiles - 1) / (kGmWarps * 8 * kGmTiles)), (unsigned)ksplit);
    const bool vec = ((uintptr_t)dt % 16 == 0) && ((uintptr_t)C16 % 16 == 0) && ((uintptr_t)logw % 16 == 0) &&
                     (ld % 2 == 0) && (ldc % 2 == 0);
    if (vec) normvar_gemm_kernel<0, 2, true><<<grid, 32 * kGmWarps, 0, st>>>(dt, genes, n, ld, C16, ldc, logw, wt, ksplit, (double*)scratch);
    else normvar_gemm_kernel<0, 2, false><<<grid, 32 * kGmWarps, 0, st>>>(dt, genes, n, ld, C16, ldc, logw, wt, ksplit, (double*)scratch);
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
    NSR_REQUIRE(genes >= 1 && n >= 1 && ld >= n && ldo >= n && ldc >= n && nc >= 1 && nc <= 16,
                "nsr_normvar_apply: bad shape genes=%lld n=%lld nc=%d (1..16 covariates)", (long long)genes,
                (long long)n, nc);
    NSR_CHECK(cudaSetDevice(ctx->device));
    const int64_t gx = (genes + kNvWarps * 16 - 1) / (kNvWarps * 16);        // 8 warps x 16 genes per CTA
    int64_t gy = (6 * (int64_t)ctx->sm_count + gx - 1) / gx;                 // >= 6 CTAs per SM in total
    const int64_t max_y = (n + 255) / 256;
    if (gy > max_y) gy = max_y;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const int64_t per = ((n + gy - 1) / gy + 15) / 16 * 16;                  // cells per split, whole steps
    gy = (n + per - 1) / per;
    const dim3 grid((unsigned)gx, (unsigned)gy);
    cudaStream_t st = (cudaStream_t)stream;
    // 16-byte accesses: rows of dt / out and log w 16-byte aligned (cell splits are multiples of 16 cells)
    const bool vec = ((uintptr_t)dt % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)logw % 16 == 0) &&
                     (ld % 2 == 0) && (ldo % 2 == 0);
#define NSR_NV_APPLY(KS_)                                                                                              \
    do {                                                                                                               \
        if (vec) normvar_apply_kernel<KS_, true><<<grid, kNvThreads, 0, st>>>(dt, genes, n, ld, dc, nc, ldc, logw, wt, coef, scale, out, ldo, per); \
        else normvar_apply_kernel<KS_, false><<<grid, kNvThreads, 0, st>>>(dt, genes, n, ld, dc, nc, ldc, logw, wt, coef, scale, out, ldo, per);     \
    } while (0)
    switch ((nc + 3) / 4) {
        case 1: NSR_NV_APPLY(1); break;
        case 2: NSR_NV_APPLY(2); break;
        case 3: NSR_NV_APPLY(3); break;
        default: NSR_NV_APPLY(4); break;
    }
#undef NSR_NV_APPLY
    NSR_CHECK(cudaGetLastError());
    return 0;
}
