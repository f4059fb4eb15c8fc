// Grouped cross-products for de(single=1) (reference association_test_2, association.py:263-390).
//
// The reference tests grouping x on its own subset of cells S_x = U + T_x (U: cells without any
// gRNA, shared by all x; T_x: cells carrying only x; driver association.py:913-916) and, for
// every x, copies dy[:, S_x] and projects the covariates out of it again.  All of its statistics
// are sums over S_x of products of covariates, x and y, and the T_x are disjoint, so
//     sum over S_x = sum over U  +  sum over T_x.
// The U part is one dense masked pass over dy (nsr_project_coef with the covariates masked to U).
// This kernel produces the T_x parts for all x at once from the columns of dy gathered in group
// order: for every group g and gene y
//     out[g][y][j]   = sum_{k in g} C[j][k] Y[y][k]     (j < nc1: covariates and a row of ones)
//     out[g][y][nc1] = sum_{k in g} Y[y][k]^2
// Y is read once (HBM-streaming, 8 B per gathered entry); sums run in a fixed order.
#include "nsr_common.cuh"
#include "pvalue.cuh"

namespace {

constexpr int kGThreads = 256;
constexpr int kGWarps = kGThreads / 32;
constexpr int kGenesPerWarp = 4;
constexpr int kChunk = 128;            // cells of a group staged per step

// NJ = covariate rows accumulated per sweep over the cells (template: no predicated-off FMAs)
template <int NJ>
__global__ void __launch_bounds__(kGThreads, NJ <= 8 ? 2 : 1)
group_stats_kernel(const double* __restrict__ Y, int64_t genes, int64_t ldy, const double* __restrict__ C, int nc1,
                   int64_t ldc, const int64_t* __restrict__ goff, double* __restrict__ out) {
    __shared__ double s_c[NJ][kChunk];
    const int g = blockIdx.y;
    const int64_t k_begin = goff[g], k_end = goff[g + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gene0 = ((int64_t)blockIdx.x * kGWarps + warp) * kGenesPerWarp;
    const int n_out = nc1 + 1;

    for (int j0 = 0; j0 < nc1 || j0 == 0; j0 += NJ) {
        const int nj = min(NJ, nc1 - j0);
        double acc[kGenesPerWarp][NJ];
        double sq[kGenesPerWarp];
#pragma unroll
        for (int q = 0; q < kGenesPerWarp; ++q) {
            sq[q] = 0.0;
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[q][j] = 0.0;
        }
        for (int64_t k0 = k_begin; k0 < k_end; k0 += kChunk) {
            const int len = (int)min((int64_t)kChunk, k_end - k0);
            __syncthreads();
            for (int idx = threadIdx.x; idx < NJ * kChunk; idx += kGThreads) {
                const int j = idx / kChunk, k = idx % kChunk;
                s_c[j][k] = (j < nj && k < len) ? C[(int64_t)(j0 + j) * ldc + k0 + k] : 0.0;
            }
            __syncthreads();
            // all loads of the chunk first (16 per lane): one resident CTA per SM cannot hide HBM latency otherwise
            double y[kChunk / 32][kGenesPerWarp];
#pragma unroll
            for (int kk = 0; kk < kChunk / 32; ++kk) {
                const int k = lane + 32 * kk;
#pragma unroll
                for (int q = 0; q < kGenesPerWarp; ++q)
                    y[kk][q] = (k < len && gene0 + q < genes) ? Y[(gene0 + q) * ldy + k0 + k] : 0.0;
            }
#pragma unroll
            for (int kk = 0; kk < kChunk / 32; ++kk) {
                const int k = lane + 32 * kk;
                if (32 * kk >= len) break;                 // warp-uniform
                double c[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) c[j] = s_c[j][k];           // read once, used for every gene of the warp
#pragma unroll
                for (int q = 0; q < kGenesPerWarp; ++q) {
                    sq[q] = fma(y[kk][q], y[kk][q], sq[q]);
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc[q][j] = fma(c[j], y[kk][q], acc[q][j]);
                }
            }
        }
        // fixed-shape butterfly: the same summation order for every launch
#pragma unroll
        for (int q = 0; q < kGenesPerWarp; ++q) {
            const int64_t gene = gene0 + q;
            if (gene >= genes) break;
            double* o = out + ((int64_t)g * genes + gene) * n_out;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                double v = acc[q][j];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0 && j < nj) o[j0 + j] = v;
            }
            if (j0 == 0) {
                double v = sq[q];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) o[nc1] = v;
            }
        }
        if (nc1 == 0) break;
    }
}

// Closed form of association_test_2 for one block of genes (association.py:352-377), fused: from
// the U parts (cu: [genes][nc + 1] = covariate-y products over U and sum_U y; yy_u[genes]) and the
// T_x parts (st: nsr_group_stats output, [groups][genes][nc + 2]) to gamma, var_y, P (and alpha):
//   cy = cu + st,  ccy = C+ cy,  Syy = yy - ccy.cy,  Sxy = sum_{T_x} y - ccy.cx,
//   var_y = Syy / ns, gamma = Sxy / (ns var_x), R2 = gamma^2 var_x / var_y, P = I_{1-R2}(dof/2, 1/2).
// One CTA per (grouping x, 256 genes); C+ (nc x nc), cx, C+ cx of the grouping sit in shared memory.
constexpr int kFinNc = 16;             // covariates handled in registers

__global__ void __launch_bounds__(256)
single1_finish_kernel(const double* __restrict__ cu, const double* __restrict__ yy_u, const double* __restrict__ st,
                      int64_t genes, int nc, const double* __restrict__ ci, const double* __restrict__ cx,
                      const double* __restrict__ ccx, const double* __restrict__ ns, const double* __restrict__ vx,
                      const double* __restrict__ dof, double* __restrict__ P, double* __restrict__ gamma,
                      double* __restrict__ vy, double* __restrict__ alpha, int64_t ld_out, int64_t col0,
                      int* __restrict__ flag) {
    __shared__ double s_ci[kFinNc * kFinNc], s_cx[kFinNc], s_ccx[kFinNc];
    __shared__ NsrPvalParams s_pv;
    const int x = blockIdx.y;
    for (int i = threadIdx.x; i < nc * nc; i += blockDim.x) s_ci[i] = ci[(int64_t)x * nc * nc + i];
    for (int i = threadIdx.x; i < nc; i += blockDim.x) { s_cx[i] = cx[(int64_t)x * nc + i]; s_ccx[i] = ccx[(int64_t)x * nc + i]; }
    if (threadIdx.x == 0) s_pv = nsr_pval_params(dof[x] * 0.5);
    __syncthreads();
    const int64_t y = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= genes) return;
    const double* u = cu + y * (nc + 1);
    const double* t = st + ((int64_t)x * genes + y) * (nc + 2);
    double cy[kFinNc];
#pragma unroll
    for (int j = 0; j < kFinNc; ++j) cy[j] = j < nc ? u[j] + t[j] : 0.0;
    double quad = 0.0, cross = 0.0, ccy[kFinNc];
#pragma unroll
    for (int i = 0; i < kFinNc; ++i) {
        double a = 0.0;
        if (i < nc) {
#pragma unroll
            for (int j = 0; j < kFinNc; ++j) if (j < nc) a = fma(s_ci[i * nc + j], cy[j], a);   // C+ symmetric
            quad = fma(a, cy[i], quad);
            cross = fma(a, s_cx[i], cross);
        }
        ccy[i] = a;
    }
    const double nsx = ns[x], vxx = vx[x];
    const double syy = (yy_u[y] + t[nc + 1]) - quad;
    const double sxy = t[nc] - cross;
    const double v_y = syy / nsx;
    const double gam = sxy / (nsx * vxx);                               // association.py:364
    const double r2 = gam * gam * vxx / v_y;                            // :368
    if (!(r2 >= 0.0 && r2 <= 1.0 + 1e-8)) atomicOr(flag, 1);           // :371
    const int64_t o = (int64_t)x * ld_out + col0 + y;
    P[o] = nsr_pvalue_r2(r2 > 1.0 ? 1.0 : r2, s_pv);
    gamma[o] = gam;
    vy[o] = v_y;
    if (alpha) {                                                        // :365-367
#pragma unroll
        for (int j = 0; j < kFinNc; ++j)
            if (j < nc) alpha[o * nc + j] = ccy[j] - gam * s_ccx[j];
    }
}

}  // namespace

extern "C" int nsr_group_stats(nsr_ctx* ctx, uintptr_t stream, const double* Y, int64_t genes, int64_t ldy,
                               const double* C, int nc1, int64_t ldc, const int64_t* goff, int n_groups,
                               double* out) {
    NSR_REQUIRE(ctx && Y && goff && out, "nsr_group_stats: null argument");
    NSR_REQUIRE(genes >= 1 && n_groups >= 1 && n_groups <= 65535 && nc1 >= 0 && nc1 <= NSR_MAX_RANK + 1 &&
                    (nc1 == 0 || C != nullptr),
                "nsr_group_stats: bad shape genes=%lld groups=%d nc1=%d", (long long)genes, n_groups, nc1);
    NSR_CHECK(cudaSetDevice(ctx->device));
    const unsigned gx = (unsigned)((genes + kGWarps * kGenesPerWarp - 1) / (kGWarps * kGenesPerWarp));
    const dim3 grid(gx, (unsigned)n_groups);
    cudaStream_t st = (cudaStream_t)stream;
#define NSR_GS(W) group_stats_kernel<W><<<grid, kGThreads, 0, st>>>(Y, genes, ldy, C, nc1, ldc, goff, out)
    if (nc1 <= 4) NSR_GS(4);
    else if (nc1 <= 8) NSR_GS(8);
    else if (nc1 <= 12) NSR_GS(12);
    else NSR_GS(16);
#undef NSR_GS
    NSR_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int nsr_single1_finish(nsr_ctx* ctx, uintptr_t stream, const double* cu, const double* yy_u,
                                  const double* st, int64_t genes, int n_groups, int nc, const double* ci,
                                  const double* cx, const double* ccx, const double* ns, const double* vx,
                                  const double* dof, double* P, double* gamma, double* vy, double* alpha,
                                  int64_t ld_out, int64_t col0, int* flag) {
    NSR_REQUIRE(ctx && cu && yy_u && st && ns && vx && dof && P && gamma && vy && flag, "nsr_single1_finish: null argument");
    NSR_REQUIRE(genes >= 1 && n_groups >= 1 && n_groups <= 65535 && nc >= 0 && nc <= kFinNc && (nc == 0 || (ci && cx && ccx)),
                "nsr_single1_finish: bad shape genes=%lld groups=%d nc=%d (nc <= %d)", (long long)genes, n_groups, nc, kFinNc);
    NSR_CHECK(cudaSetDevice(ctx->device));
    const dim3 grid((unsigned)((genes + 255) / 256), (unsigned)n_groups);
    single1_finish_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cu, yy_u, st, genes, nc, ci, cx, ccx, ns, vx, dof, P, gamma,
                                                                    vy, alpha, ld_out, col0, flag);
    NSR_CHECK(cudaGetLastError());
    return 0;
}
