// Covariate-basis helpers: the two products with the (covariates x cells) matrix that the host
// needs to turn `dc` into an orthonormal basis Qt of its row space (association.py:899-903 computes
// dci, dcr = inv_rank(dc dc^T); here the host factorises the same nc x nc Gram matrix and the
// device applies the resulting small matrix).  nc <= NSR_MAX_RANK rows, so both kernels are
// HBM-streaming over a few MB; partial sums are combined in a fixed order (deterministic).
#include "nsr_common.cuh"

namespace {

constexpr int kGramChunk = 1024;   // cells per CTA
constexpr int kGramTile = 32;      // cells staged per step
constexpr int kGramThreads = 256;

__global__ void __launch_bounds__(kGramThreads)
cov_gram_partial_kernel(const double* __restrict__ C, int nc, int64_t n, int64_t ldc, double* __restrict__ partial) {
    __shared__ double s[NSR_MAX_RANK][kGramTile + 1];
    const int64_t c0 = (int64_t)blockIdx.x * kGramChunk;
    const int64_t c1 = min(n, c0 + kGramChunk);
    const int n_entry = nc * nc;
    constexpr int kPer = NSR_MAX_RANK * NSR_MAX_RANK / kGramThreads;
    double acc[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) acc[q] = 0.0;
    for (int64_t k0 = c0; k0 < c1; k0 += kGramTile) {
        for (int idx = threadIdx.x; idx < nc * kGramTile; idx += kGramThreads) {
            const int r = idx / kGramTile, k = idx % kGramTile;
            s[r][k] = (k0 + k < c1) ? C[(int64_t)r * ldc + k0 + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int e = threadIdx.x + q * kGramThreads;
            if (e < n_entry) {
                const int i = e / nc, j = e % nc;
                double a = acc[q];
#pragma unroll 8
                for (int k = 0; k < kGramTile; ++k) a = fma(s[i][k], s[j][k], a);
                acc[q] = a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int e = threadIdx.x + q * kGramThreads;
        if (e < n_entry) partial[(int64_t)blockIdx.x * n_entry + e] = acc[q];
    }
}

__global__ void cov_gram_reduce_kernel(const double* __restrict__ partial, int n_entry, int n_part, double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entry) return;
    double a = 0.0;
    for (int p = 0; p < n_part; ++p) a += partial[(int64_t)p * n_entry + e];
    G[e] = a;
}

__global__ void __launch_bounds__(256)
cov_apply_kernel(const double* __restrict__ M, int rank, int nc, const double* __restrict__ C, int64_t n,
                 int64_t ldc, double* __restrict__ Q, int64_t ldq) {
    extern __shared__ double s_m[];
    for (int idx = threadIdx.x; idx < rank * nc; idx += blockDim.x) s_m[idx] = M[idx];
    __syncthreads();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    for (int r = 0; r < rank; ++r) {
        double a = 0.0;
        for (int c = 0; c < nc; ++c) a = fma(s_m[r * nc + c], C[(int64_t)c * ldc + k], a);
        Q[(int64_t)r * ldq + k] = a;
    }
}

}  // namespace

extern "C" int nsr_cov_gram(nsr_ctx* ctx, uintptr_t stream, const double* C, int nc, int64_t n, int64_t ldc,
                            double* G) {
    NSR_REQUIRE(ctx && C && G, "nsr_cov_gram: null argument");
    NSR_REQUIRE(nc >= 1 && nc <= NSR_MAX_RANK && n >= 1 && ldc >= n, "nsr_cov_gram: bad shape nc=%d n=%lld", nc,
                (long long)n);
    NSR_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n_part = (int)((n + kGramChunk - 1) / kGramChunk);
    void* scratch = nullptr;
    if (nsr_scratch(ctx, (size_t)n_part * nc * nc * sizeof(double), &scratch)) return 1;
    cov_gram_partial_kernel<<<n_part, kGramThreads, 0, st>>>(C, nc, n, ldc, (double*)scratch);
    cov_gram_reduce_kernel<<<(nc * nc + 127) / 128, 128, 0, st>>>((const double*)scratch, nc * nc, n_part, G);
    NSR_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int nsr_cov_apply(nsr_ctx* ctx, uintptr_t stream, const double* M, int rank, int nc, const double* C,
                             int64_t n, int64_t ldc, double* Q, int64_t ldq) {
    NSR_REQUIRE(ctx && M && C && Q, "nsr_cov_apply: null argument");
    NSR_REQUIRE(rank >= 1 && rank <= NSR_MAX_RANK && nc >= 1 && nc <= NSR_MAX_RANK && n >= 1 && ldc >= n && ldq >= n,
                "nsr_cov_apply: bad shape rank=%d nc=%d n=%lld", rank, nc, (long long)n);
    NSR_CHECK(cudaSetDevice(ctx->device));
    cov_apply_kernel<<<(unsigned)((n + 255) / 256), 256, (size_t)rank * nc * sizeof(double), (cudaStream_t)stream>>>(
        M, rank, nc, C, n, ldc, Q, ldq);
    NSR_CHECK(cudaGetLastError());
    return 0;
}
