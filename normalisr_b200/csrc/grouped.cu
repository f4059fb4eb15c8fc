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
