// de(single=4) on the device: the leave-one-out regression of association_test_4
// (reference src/normalisr/association.py:421-576) from the Gram matrices of the covariate-residualised
// groupings, without a per-grouping pseudo-inverse.
//
// With Gxx = Rx Rx^T (nx x nx), Gxy = Rx Ry^T (nx x ny), yy = rowsum(Ry^2) and K = Gxx^-1, testing
// grouping x with all other groupings as nuisance regressors is the multiple regression of y on all
// rows of Rx, read off K (Schur complements; the reference's lines in brackets):
//     w    = K Gxy                                       full-regression coefficients
//     dxx  = 1 / (n K_xx)                                [:539-540]  x given the others
//     dxy  = w_xy / (n K_xx)                             [:543-544]
//     dyy  = (yy_y - sum_x Gxy_xy w_xy + w_xy^2 / K_xx) / n          [:541-542]  y given the others
//     gamma = dxy / dxx,  R2 = dxy^2 / (dxx dyy),  P = I_{1-R2}((n - 1 - rank - dimreduce) / 2, 1/2)
// rank = nx - 1 + rank(C) for every x when Gxx is positive definite at the reference's tolerance.
//
// Kernels (all float64, fixed summation orders):
//   gemm_f64_kernel        C = op(A) op(B), 64 x 64 tiles, 4 x 4 per thread, optional split over K
//   chol_panel / chol_update  blocked right-looking Cholesky, 32-column panels
//   tri_inverse_(smem_)kernel  L^-1, one warp per column (the column in shared memory up to n = 3,200)
//   de4_colsum / de4_finish   the closed form above + the P-value
#include "epilogue.cuh"

namespace {

constexpr int kTile = 64;
constexpr int kKT = 16;

// C[m][n] (ldc) = sum_k A(m,k) B(k,n);  A(m,k) = A[m * sam + k * sak], B(k,n) = B[k * sbk + n * sbn].
// grid (tiles_n, tiles_m, ksplit): with ksplit > 1 the slice z writes its partial sum to
// C + z * split_stride and gemm_reduce_kernel adds the slices in order.
__global__ void __launch_bounds__(256)
gemm_f64_kernel(const double* __restrict__ A, int64_t sam, int64_t sak, const double* __restrict__ B, int64_t sbk,
                int64_t sbn, double* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t split_stride,
                int lower_only) {
    __shared__ double sA[kKT][kTile + 1];
    __shared__ double sB[kKT][kTile + 1];
    const int64_t m0 = (int64_t)blockIdx.y * kTile, n0 = (int64_t)blockIdx.x * kTile;
    if (lower_only && n0 > m0) return;                        // symmetric result: tiles on or below the diagonal
    const int64_t kper = (K + gridDim.z - 1) / gridDim.z;
    const int64_t kb = kper * blockIdx.z, ke = (kb + kper < K) ? kb + kper : K;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    // loaders: when the K stride is 1 consecutive threads walk k (coalesced), otherwise they walk m / n
    const bool a_kfast = sak == 1, b_kfast = sbk == 1;
    for (int64_t k0 = kb; k0 < ke; k0 += kKT) {
        for (int e = threadIdx.x; e < kKT * kTile; e += 256) {
            int kk, mm;
            if (a_kfast) { kk = e % kKT; mm = e / kKT; } else { mm = e % kTile; kk = e / kTile; }
            const int64_t m = m0 + mm, k = k0 + kk;
            sA[kk][mm] = (m < M && k < ke) ? A[m * sam + k * sak] : 0.0;
            int nn;
            if (b_kfast) { kk = e % kKT; nn = e / kKT; } else { nn = e % kTile; kk = e / kTile; }
            const int64_t nc = n0 + nn, k2 = k0 + kk;
            sB[kk][nn] = (nc < N && k2 < ke) ? B[k2 * sbk + nc * sbn] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kKT; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    double* Cz = C + (int64_t)blockIdx.z * split_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t nc = n0 + tx + 16 * j;
            if (nc < N) Cz[m * ldc + nc] = acc[i][j];
        }
    }
}

// out[m][n] = alpha * sum_z part[z][m][n] + beta * sub[m][n]  (fixed order); lower_only: the strict upper
// triangle is mirrored from the lower one
__global__ void gemm_reduce_kernel(const double* __restrict__ part, int ksplit, int64_t split_stride, int64_t ldp,
                                   double* __restrict__ out, int64_t ldo, int64_t M, int64_t N, double alpha,
                                   const double* __restrict__ sub, int64_t lds, double beta, int lower_only) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const int64_t m = i / N, nc = i % N;
    const int64_t pm = (lower_only && nc > m) ? nc : m, pn = (lower_only && nc > m) ? m : nc;
    double s = 0.0;
    for (int z = 0; z < ksplit; ++z) s += part[(int64_t)z * split_stride + pm * ldp + pn];
    s *= alpha;
    if (sub) s = fma(beta, sub[pm * lds + pn], s);
    out[m * ldo + nc] = s;
}

// ---- Cholesky, lower, in place, panels of kNB columns -------------------------------------------
constexpr int kNB = 32;
constexpr int kPanelRows = 128;      // rows of the panel one CTA solves (128 threads)

// Factor the diagonal block A[k0:k0+nb][k0:k0+nb] in place: ONE warp, lane r holds row r in registers and
// the pivot column travels by shuffle - no shared memory, no block barrier (the barrier-per-step version spent
// 28 us per 32 x 32 block, a quarter of the whole solve at 300 groupings).  A pivot <= tol * max diag (or not
// finite) raises *status |= 1 and is replaced by 1 so that nothing overflows.
__global__ void __launch_bounds__(32)
chol_diag_kernel(double* __restrict__ A, int64_t lda, int64_t k0, int nb, double tol,
                 const double* __restrict__ diag_max, int* __restrict__ status) {
    const int r = threadIdx.x;
    const double piv_floor = tol * *diag_max;
    double row[kNB];
#pragma unroll
    for (int c = 0; c < kNB; ++c) row[c] = (r < nb && c <= r) ? A[(k0 + r) * lda + k0 + c] : 0.0;
#pragma unroll
    for (int j = 0; j < kNB; ++j) {
        if (j < nb) {                                        // nb is warp-uniform
            double d = __shfl_sync(0xffffffffu, row[j], j);  // pivot A[j][j] (already updated)
            if (!(d > piv_floor) || !isfinite(d)) {
                if (r == 0) atomicOr(status, 1);
                d = 1.0;
            }
            const double dj = sqrt(d);
            if (r == j) row[j] = dj;
            else if (r > j) row[j] /= dj;
            // trailing update: A[r][c] -= L[r][j] L[c][j] for j < c <= r; L[c][j] comes from lane c
            const double lrj = row[j];
#pragma unroll
            for (int c = j + 1; c < kNB; ++c) {
                const double lcj = __shfl_sync(0xffffffffu, lrj, c);
                if (c <= r) row[c] = fma(-lrj, lcj, row[c]);
            }
        }
    }
    if (r < nb) {
#pragma unroll
        for (int c = 0; c < kNB; ++c)
            if (c <= r) A[(k0 + r) * lda + k0 + c] = row[c];
    }
}

// L21 rows [k0 + nb + b * kPanelRows, ...) = A21 L11^-T, one row per thread, L11 (factored) in shared memory
__global__ void __launch_bounds__(kPanelRows)
chol_panel_kernel(double* __restrict__ A, int64_t n, int64_t lda, int64_t k0, int nb) {
    __shared__ double L[kNB][kNB + 1];
    const int t = threadIdx.x;
    for (int e = t; e < nb * nb; e += kPanelRows) {
        const int r = e / nb, c = e % nb;
        L[r][c] = c <= r ? A[(k0 + r) * lda + k0 + c] : 0.0;
    }
    __syncthreads();
    const int64_t row = k0 + nb + (int64_t)blockIdx.x * kPanelRows + t;
    if (row < n) {
        double x[kNB];
#pragma unroll
        for (int j = 0; j < kNB; ++j) {
            if (j < nb) {
                double s = A[row * lda + k0 + j];
#pragma unroll
                for (int i = 0; i < kNB; ++i)
                    if (i < j) s = fma(-x[i], L[j][i], s);
                x[j] = s / L[j][j];
            }
        }
#pragma unroll
        for (int j = 0; j < kNB; ++j)
            if (j < nb) A[row * lda + k0 + j] = x[j];
    }
}

// A22 -= L21 L21^T on the lower-triangular 64 x 64 tiles of the trailing matrix (K = nb <= 32)
__global__ void __launch_bounds__(256)
chol_update_kernel(double* __restrict__ A, int64_t n, int64_t lda, int64_t k0, int nb) {
    const int64_t base = k0 + nb;
    const int64_t m0 = base + (int64_t)blockIdx.y * kTile, n0 = base + (int64_t)blockIdx.x * kTile;
    if (n0 > m0 || m0 >= n) return;
    __shared__ double sR[kNB][kTile + 1], sC[kNB][kTile + 1];
    for (int e = threadIdx.x; e < nb * kTile; e += 256) {
        const int kk = e % nb, mm = e / nb;
        sR[kk][mm] = (m0 + mm < n) ? A[(m0 + mm) * lda + k0 + kk] : 0.0;
        sC[kk][mm] = (n0 + mm < n) ? A[(n0 + mm) * lda + k0 + kk] : 0.0;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int kk = 0; kk < nb; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sR[kk][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sC[kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty + 16 * i;
        if (m >= n) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = n0 + tx + 16 * j;
            if (c <= m) A[m * lda + c] -= acc[i][j];
        }
    }
}

// Linv = L^-1 (lower triangular, dense n x n output with zeros above the diagonal): one warp per column j
// solves L x = e_j by forward substitution; the dot products are split over the lanes and combined with a
// fixed shuffle tree.
__global__ void __launch_bounds__(256)
tri_inverse_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, double* __restrict__ Linv, int64_t ldi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + warp;
    if (j >= n) return;
    // x lives in the output column (global memory, L2 resident); lanes read what earlier steps wrote
    for (int64_t i = lane; i < j; i += 32) Linv[i * ldi + j] = 0.0;
    if (lane == 0) Linv[j * ldi + j] = 1.0 / L[j * ldl + j];
    __syncwarp();
    for (int64_t i = j + 1; i < n; ++i) {
        double s = 0.0;
        for (int64_t k = j + lane; k < i; k += 32) s = fma(L[i * ldl + k], Linv[k * ldi + j], s);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if (lane == 0) Linv[i * ldi + j] = -s / L[i * ldl + i];
        __syncwarp();
    }
}

// The same substitution with the column being built kept in shared memory (8 columns x n doubles per CTA): the
// dependent chain of a step is a shared-memory read, a shuffle tree and a division instead of a round trip through
// L2 for what the previous step wrote.  Same lane partition, same tree, same arithmetic: identical bits.
__global__ void __launch_bounds__(256)
tri_inverse_smem_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, double* __restrict__ Linv, int64_t ldi) {
    extern __shared__ double s_cols[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + warp;
    if (j >= n) return;
    double* x = s_cols + (int64_t)warp * n;
    for (int64_t i = lane; i < j; i += 32) Linv[i * ldi + j] = 0.0;
    if (lane == 0) x[j] = 1.0 / L[j * ldl + j];
    __syncwarp();
    for (int64_t i = j + 1; i < n; ++i) {
        const double diag = L[i * ldl + i];                       // (not on the dependent chain)
        double s = 0.0;
        for (int64_t k = j + lane; k < i; k += 32) s = fma(L[i * ldl + k], x[k], s);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if (lane == 0) x[i] = -s / diag;
        __syncwarp();
    }
    for (int64_t i = j + lane; i < n; i += 32) Linv[i * ldi + j] = x[i];
}

// ... and with the next row of L already on its way while a step's chain (shared-memory reads, shuffle tree,
// division) runs: the row segments live in registers, MAXK per lane (n <= 32 MAXK).  Terms beyond the row's length
// are fma(0, 0, s) = s, so the sums are the same bits as above.
template <int MAXK>
__global__ void __launch_bounds__(256)
tri_inverse_pipe_kernel(const double* __restrict__ L, int64_t n, int64_t ldl, double* __restrict__ Linv, int64_t ldi) {
    extern __shared__ double s_cols[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + warp;
    if (j >= n) return;
    double* x = s_cols + (int64_t)warp * n;
    for (int64_t i = lane; i < j; i += 32) Linv[i * ldi + j] = 0.0;
    if (lane == 0) x[j] = 1.0 / L[j * ldl + j];
    __syncwarp();
    double cur[MAXK], nxt[MAXK], dcur = 1.0, dnxt = 1.0;
    auto load_row = [&](int64_t i, double (&v)[MAXK], double& dg) {
#pragma unroll
        for (int m = 0; m < MAXK; ++m) {
            const int64_t k = j + lane + 32 * m;
            v[m] = k < i ? L[i * ldl + k] : 0.0;
        }
        dg = L[i * ldl + i];
    };
    if (j + 1 < n) load_row(j + 1, cur, dcur);
    for (int64_t i = j + 1; i < n; ++i) {
        if (i + 1 < n) load_row(i + 1, nxt, dnxt);
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < MAXK; ++m) {
            const int64_t k = j + lane + 32 * m;
            s = fma(cur[m], k < i ? x[k] : 0.0, s);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if (lane == 0) x[i] = -s / dcur;
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MAXK; ++m) cur[m] = nxt[m];
        dcur = dnxt;
    }
    for (int64_t i = j + lane; i < n; i += 32) Linv[i * ldi + j] = x[i];
}

__global__ void diag_kernel(const double* __restrict__ K, int64_t n, int64_t ldk, double* __restrict__ kd) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) kd[i] = K[i * ldk + i];
}

// *out = max_i G[i][i]  (one block)
__global__ void diag_max_kernel(const double* __restrict__ G, int64_t n, int64_t ldg, double* __restrict__ out) {
    __shared__ double sm[256];
    double m = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) m = fmax(m, G[i * ldg + i]);
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int h = 128; h >= 1; h >>= 1) {
        if ((int)threadIdx.x < h) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + h]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

// qf[y] = sum_x Gxy[x][y] w[x][y].  A CTA owns 32 columns; its 8 warps take the rows x = warp, warp + 8, ...
// (coalesced 256-byte row segments) and their partial sums are combined in warp order: fixed summation order
// for a given nx, and enough CTAs (ny / 32) to stream the two matrices at memory speed.
__global__ void __launch_bounds__(256)
de4_colsum_kernel(const double* __restrict__ Gxy, int64_t ldg, const double* __restrict__ w, int64_t ldw,
                  int64_t nx, int64_t ny, double* __restrict__ qf) {
    __shared__ double part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t y = (int64_t)blockIdx.x * 32 + lane;
    double s = 0.0;
    if (y < ny)
        for (int64_t x = warp; x < nx; x += 8) s = fma(Gxy[x * ldg + y], w[x * ldw + y], s);
    part[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && y < ny) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][lane];
        qf[y] = t;
    }
}

__global__ void de4_finish_kernel(const double* __restrict__ w, int64_t ldw, const double* __restrict__ kd,
                                  const double* __restrict__ yy, const double* __restrict__ qf, int64_t nx, int64_t ny,
                                  double n_cells, int return_dot, NsrPvalParams pv, double* __restrict__ P,
                                  double* __restrict__ out2, double* __restrict__ vary, int64_t ldo,
                                  double* __restrict__ varx, int* __restrict__ status) {
    const int64_t y = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t x = blockIdx.y;
    if (y >= ny) return;
    const double k = kd[x];
    double dxx = 1.0 / (n_cells * k);
    if (dxx == 0.0) dxx = 1.0;                                   // association.py:545-547
    const double wv = w[x * ldw + y];
    const double dxy = wv / k / n_cells;
    const double dyy = (yy[y] - qf[y] + wv * wv / k) / n_cells;
    const double gamma = dxy / dxx;
    const double r2 = dxy * dxy / (dxx * dyy);
    if (!(r2 >= 0.0 && r2 <= 1.0 + 1e-8) || !isfinite(gamma) || !(dyy >= 0.0)) atomicOr(status, 2);   // :557, :565-568
    P[x * ldo + y] = nsr_pvalue_r2(r2, pv);
    out2[x * ldo + y] = return_dot ? gamma * dxx : gamma;
    vary[x * ldo + y] = dyy;
    if (y == 0) varx[x] = dxx;
}

int launch_gemm(nsr_ctx* ctx, cudaStream_t st, const double* A, int64_t sam, int64_t sak, const double* B, int64_t sbk,
                int64_t sbn, double* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int lower_only, double alpha,
                const double* sub, int64_t lds, double beta, double* work, size_t work_doubles) {
    const dim3 tiles((unsigned)((N + kTile - 1) / kTile), (unsigned)((M + kTile - 1) / kTile));
    int ksplit = 1;
    const int64_t n_tiles = (int64_t)tiles.x * tiles.y;
    if (n_tiles < 2 * ctx->sm_count && K >= 4096) {
        ksplit = (int)((4 * ctx->sm_count + n_tiles - 1) / n_tiles);
        if (ksplit > (int)(K / 1024)) ksplit = (int)(K / 1024);
        while (ksplit > 1 && (size_t)ksplit * M * N > work_doubles) --ksplit;
        if (ksplit < 1) ksplit = 1;
    }
    if (ksplit == 1 && alpha == 1.0 && sub == nullptr && !lower_only) {
        gemm_f64_kernel<<<tiles, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, 0, 0);
    } else {
        NSR_REQUIRE((size_t)ksplit * M * N <= work_doubles, "de4: gemm workspace too small");
        gemm_f64_kernel<<<dim3(tiles.x, tiles.y, (unsigned)ksplit), 256, 0, st>>>(A, sam, sak, B, sbk, sbn, work, N, M, N, K,
                                                                              M * N, lower_only);
        gemm_reduce_kernel<<<(unsigned)((M * N + 255) / 256), 256, 0, st>>>(work, ksplit, M * N, N, C, ldc, M, N, alpha, sub,
                                                                         lds, beta, lower_only);
    }
    NSR_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace

// G = X X^T - Cx Cx^T for a (rows x n) float64 matrix X and its (rows x rank) covariate coefficients
// Cx = X Qt^T (from nsr_residualize with keep-coef / nsr_project_coef): the float64 Gram matrix of the
// residualised rows, formed like the reference forms prod1 tiles (association.py:393-418) from raw rows.
extern "C" int nsr_gram_f64(nsr_ctx* ctx, uintptr_t stream, const double* X, int64_t rows, int64_t n, int64_t ldx,
                            const double* coef, int rank, double* G, int64_t ldg) {
    NSR_REQUIRE(ctx && X && G && rows > 0 && n > 0 && ldx >= n && ldg >= rows && rank >= 0 && (rank == 0 || coef),
                "nsr_gram_f64: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    NSR_CHECK(cudaSetDevice(ctx->device));
    const size_t work_doubles = (size_t)rows * rows * 64 + (size_t)rows * rows;
    void* scratch = nullptr;
    if (nsr_scratch(ctx, work_doubles * sizeof(double), &scratch)) return 1;
    double* work = (double*)scratch;
    double* cc = work + (size_t)rows * rows * 64;
    if (rank > 0) {
        if (launch_gemm(ctx, st, coef, rank, 1, coef, 1, rank, cc, rows, rows, rows, rank, 0, 1.0, nullptr, 0, 0.0, nullptr, 0))
            return 1;
    }
    // X X^T over the cells, split over K; the reduce subtracts Cx Cx^T and mirrors the lower triangle
    const dim3 tiles((unsigned)((rows + kTile - 1) / kTile), (unsigned)((rows + kTile - 1) / kTile));
    int ksplit = (int)((4 * ctx->sm_count + (int64_t)tiles.x * (tiles.x + 1) / 2 - 1) / ((int64_t)tiles.x * (tiles.x + 1) / 2));
    if (ksplit > 64) ksplit = 64;
    if ((int64_t)ksplit > (n + 1023) / 1024) ksplit = (int)((n + 1023) / 1024);
    if (ksplit < 1) ksplit = 1;
    gemm_f64_kernel<<<dim3(tiles.x, tiles.y, (unsigned)ksplit), 256, 0, st>>>(X, ldx, 1, X, 1, ldx, work, rows, rows, rows, n,
                                                                          rows * rows, 1);
    gemm_reduce_kernel<<<(unsigned)((rows * rows + 255) / 256), 256, 0, st>>>(work, ksplit, rows * rows, rows, G, ldg, rows, rows,
                                                                           1.0, rank > 0 ? cc : nullptr, rows, -1.0, 1);
    NSR_CHECK(cudaGetLastError());
    return 0;
}

// G -= Cx Cx^T (in place, symmetric): turns a Gram matrix of RAW rows (e.g. exact integer sums from
// nsr_contract_ab in NSR_MODE_RAW on nsr_residualize_exact planes) into that of the residualised rows.
namespace {
__global__ void sub_sym_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ G, int64_t ldg,
                               const double* __restrict__ cc, int64_t rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * rows) return;
    const int64_t m = i / rows, c = i % rows;
    // symmetric in exact arithmetic; take both from the lower triangle so the result is exactly symmetric
    const int64_t pm = c > m ? c : m, pn = c > m ? m : c;
    G[m * ldg + c] = 0.5 * (src[pm * lds + pn] + src[pn * lds + pm]) - cc[pm * rows + pn];
}
}  // namespace
extern "C" int nsr_gram_correct(nsr_ctx* ctx, uintptr_t stream, double* G, int64_t rows, int64_t ldg, const double* coef,
                                int rank) {
    NSR_REQUIRE(ctx && G && rows > 0 && ldg >= rows && rank >= 0 && (rank == 0 || coef), "nsr_gram_correct: bad arguments");
    if (rank == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    NSR_CHECK(cudaSetDevice(ctx->device));
    void* scratch = nullptr;
    if (nsr_scratch(ctx, (size_t)rows * rows * 2 * sizeof(double), &scratch)) return 1;
    double* cc = (double*)scratch;
    double* tmp = cc + rows * rows;
    if (launch_gemm(ctx, st, coef, rank, 1, coef, 1, rank, cc, rows, rows, rows, rank, 0, 1.0, nullptr, 0, 0.0, nullptr, 0)) return 1;
    // out of place (tmp -> G): every output element reads two input elements
    NSR_CHECK(cudaMemcpy2DAsync(tmp, rows * sizeof(double), G, ldg * sizeof(double), rows * sizeof(double), rows,
                                cudaMemcpyDeviceToDevice, st));
    sub_sym_kernel<<<(unsigned)((rows * rows + 255) / 256), 256, 0, st>>>(tmp, rows, G, ldg, cc, rows);
    NSR_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int nsr_de4_solve(nsr_ctx* ctx, uintptr_t stream, double* Gxx, int nx, const double* Gxy, int64_t ny,
                             int64_t ld_xy, const double* yy, int64_t n_cells, int rank_c, int dimreduce, double tol,
                             int return_dot, double* P, double* out2, double* vary, int64_t ld_out, double* varx,
                             double* w, int64_t ld_w, int* status) {
    NSR_REQUIRE(ctx && Gxx && Gxy && yy && P && out2 && vary && varx && w && status, "nsr_de4_solve: null argument");
    NSR_REQUIRE(nx > 0 && ny > 0 && ld_xy >= ny && ld_out >= ny && ld_w >= ny && n_cells > 0 && tol > 0.0,
                "nsr_de4_solve: bad shape nx=%d ny=%lld", nx, (long long)ny);
    const double dof = (double)n_cells - 1.0 - (double)(nx - 1 + rank_c) - (double)dimreduce;
    NSR_REQUIRE(dof > 0.0, "nsr_de4_solve: no degrees of freedom left");
    cudaStream_t st = (cudaStream_t)stream;
    NSR_CHECK(cudaSetDevice(ctx->device));
    const int64_t n = nx;
    // scratch: Linv | K | kd | qf | diagonal maximum | gemm workspace
    const size_t n_work = (size_t)n * n * 8;
    const size_t need = ((size_t)(2 * n * n + n + ny + 2) + n_work) * sizeof(double);
    void* scratch = nullptr;
    if (nsr_scratch(ctx, need, &scratch)) return 1;
    double* Linv = (double*)scratch;
    double* K = Linv + n * n;
    double* kd = K + n * n;
    double* qf = kd + n;
    double* dmax = qf + ny;
    double* work = dmax + 2;
    // Pivot floor tol * max diag: a Cholesky pivot is the squared norm of a grouping's component
    // orthogonal to the earlier groupings and is bounded below by the smallest eigenvalue, so a Gram
    // matrix the reference's rule (eigenvalues >= tol * largest, association.py:77) calls full rank
    // always factorises; *status |= 1 otherwise (rank-deficient groupings: the caller's other branch).
    diag_max_kernel<<<1, 256, 0, st>>>(Gxx, n, n, dmax);
    for (int64_t k0 = 0; k0 < n; k0 += kNB) {
        const int nb = (int)(n - k0 < kNB ? n - k0 : kNB);
        const int64_t below = n - k0 - nb;
        chol_diag_kernel<<<1, 32, 0, st>>>(Gxx, n, k0, nb, tol, dmax, status);
        if (below > 0) {
            chol_panel_kernel<<<(unsigned)((below + kPanelRows - 1) / kPanelRows), kPanelRows, 0, st>>>(Gxx, n, n, k0, nb);
            const unsigned tb = (unsigned)((below + kTile - 1) / kTile);
            chol_update_kernel<<<dim3(tb, tb), 256, 0, st>>>(Gxx, n, n, k0, nb);
        }
    }
    if (8 * n * (int64_t)sizeof(double) <= 200 * 1024) {
        const int smem = (int)(8 * n * sizeof(double));
        const unsigned blocks = (unsigned)((n + 7) / 8);
        if (n <= 320) {
            tri_inverse_pipe_kernel<10><<<blocks, 256, smem, st>>>(Gxx, n, n, Linv, n);
        } else if (n <= 640) {
            NSR_CHECK(cudaFuncSetAttribute(tri_inverse_pipe_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            tri_inverse_pipe_kernel<20><<<blocks, 256, smem, st>>>(Gxx, n, n, Linv, n);
        } else {
            NSR_CHECK(cudaFuncSetAttribute(tri_inverse_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            tri_inverse_smem_kernel<<<blocks, 256, smem, st>>>(Gxx, n, n, Linv, n);
        }
    } else {
        tri_inverse_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(Gxx, n, n, Linv, n);
    }
    NSR_CHECK(cudaGetLastError());
    // K = Linv^T Linv (symmetric), w = K Gxy
    if (launch_gemm(ctx, st, Linv, 1, n, Linv, n, 1, K, n, n, n, n, 0, 1.0, nullptr, 0, 0.0, work, n_work)) return 1;
    diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(K, n, n, kd);
    if (launch_gemm(ctx, st, K, n, 1, Gxy, ld_xy, 1, w, ld_w, n, ny, n, 0, 1.0, nullptr, 0, 0.0, work, n_work)) return 1;
    de4_colsum_kernel<<<(unsigned)((ny + 31) / 32), 256, 0, st>>>(Gxy, ld_xy, w, ld_w, n, ny, qf);
    const NsrPvalParams pv = nsr_pval_params(dof / 2.0);
    de4_finish_kernel<<<dim3((unsigned)((ny + 255) / 256), (unsigned)n), 256, 0, st>>>(
        w, ld_w, kd, yy, qf, n, ny, (double)n_cells, return_dot, pv, P, out2, vary, ld_out, varx, status);
    NSR_CHECK(cudaGetLastError());
    return 0;
}
