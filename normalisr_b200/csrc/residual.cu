// Covariate projection + quantisation ("residualize"): the HBM-streaming half of the path.
//
// Reference: association.py:226-233
//     ccx = dci @ (dc @ dx.T);  dx1 = dx - ccx.T @ dc;  var = mean(dx1**2), 0 -> 1
// Here the host supplies Qt, an orthonormal basis (rank x n) of the row space of dc, so
//     coef = X Qt^T  (pass A),   z = X - coef Qt  (passes B1/B2).
// B1 measures var and max|z'| per row, B2 re-derives z' and writes the int8 digit planes
// the tensor-core contraction reads.  z' is z after a sign-randomised 128-point
// Walsh-Hadamard transform along cells; it is orthonormal, so every inner product over
// cells is unchanged, but rows become near-Gaussian and a per-row fixed-point scale then
// costs no precision even for genes expressed in a handful of cells.
//
// Layout: each warp owns one 128-cell block at a time; lane l holds cells l, l+32, l+64,
// l+96 of the block, so every global load is a fully coalesced 256 B row segment.
// The covariate block (Qt) is held in registers and reused across the CTA's rows.
#include "nsr_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsA = 4;   // rows per CTA in the coefficient kernel
constexpr int kRowsB = 8;   // rows per CTA in the residual kernels
constexpr double kHadScale = 0.088388347648318440550;   // 1/sqrt(128)

__device__ __forceinline__ bool cell_flip(uint64_t k) {
    uint32_t h = (uint32_t)k ^ (uint32_t)(k >> 32) * 0x9e3779b9u;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h & 1u;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

// ---- pass A: partial[ks][row][c] = sum over the CTA's cells of X[row][k] Qt[c][k] ------
template <int CB>
__global__ void __launch_bounds__(kThreads)
coef_kernel(const double* __restrict__ X, int64_t rows, int64_t n, int64_t ldx,
            const double* __restrict__ Qt, int rank, int64_t ldq, int c0, int nblk, int ksplit,
            double* __restrict__ partial) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * kRowsA;
    const int ks = blockIdx.y;
    const int b_begin = (int)((int64_t)nblk * ks / ksplit);
    const int b_end = (int)((int64_t)nblk * (ks + 1) / ksplit);

    double acc[kRowsA][CB];
#pragma unroll
    for (int r = 0; r < kRowsA; ++r)
#pragma unroll
        for (int c = 0; c < CB; ++c) acc[r][c] = 0.0;

    for (int blk = b_begin + warp; blk < b_end; blk += kWarps) {
        const int64_t k0 = (int64_t)blk * 128 + lane;
        double q[CB][4];
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + 32 * j;
                q[c][j] = (c0 + c < rank && k < n) ? __ldg(Qt + (int64_t)(c0 + c) * ldq + k) : 0.0;
            }
#pragma unroll
        for (int r = 0; r < kRowsA; ++r) {
            if (row0 + r < rows) {
                const double* xr = X + (row0 + r) * ldx;
                double x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t k = k0 + 32 * j;
                    x[j] = (k < n) ? __ldg(xr + k) : 0.0;
                }
#pragma unroll
                for (int c = 0; c < CB; ++c)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[r][c] = fma(x[j], q[c][j], acc[r][c]);
            }
        }
    }
    __shared__ double red[kWarps][kRowsA][CB];
#pragma unroll
    for (int r = 0; r < kRowsA; ++r)
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            const double v = warp_sum(acc[r][c]);
            if (lane == 0) red[warp][r][c] = v;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < kRowsA * CB; i += kThreads) {
        const int r = i / CB, c = i % CB;
        double s = 0.0;
        for (int w = 0; w < kWarps; ++w) s += red[w][r][c];
        if (row0 + r < rows && c0 + c < rank)
            partial[((int64_t)ks * rows + row0 + r) * rank + c0 + c] = s;
    }
}

__global__ void coef_reduce_kernel(const double* __restrict__ partial, int64_t count, int ksplit,
                                   double* __restrict__ coef) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) s += partial[(int64_t)ks * count + i];   // fixed order
    coef[i] = s;
}

// ---- passes B1 / B2 ---------------------------------------------------------------------
template <bool WRITE, bool HAD>
__global__ void __launch_bounds__(kThreads)
residual_kernel(const double* __restrict__ X, int64_t rows, int64_t n, int64_t ldx,
                const double* __restrict__ Qt, int rank, int64_t ldq,
                const double* __restrict__ coef, int nblk, int ksplit, uint64_t cell_offset,
                double* __restrict__ part_sumsq, double* __restrict__ part_amax,
                const double* __restrict__ inv_quantum, int n_slices, double vmax,
                int8_t* __restrict__ slices, int64_t rows_alloc, int64_t n_pad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * kRowsB;
    const int ks = blockIdx.y;
    const int b_begin = (int)((int64_t)nblk * ks / ksplit);
    const int b_end = (int)((int64_t)nblk * (ks + 1) / ksplit);
    const int rank4 = (rank + 3) & ~3;

    __shared__ double s_coef[kRowsB][NSR_MAX_RANK];
    __shared__ double s_red[2][kWarps][kRowsB];
    for (int i = threadIdx.x; i < kRowsB * NSR_MAX_RANK; i += kThreads) {
        const int r = i / NSR_MAX_RANK, c = i % NSR_MAX_RANK;
        s_coef[r][c] = (row0 + r < rows && c < rank) ? coef[(row0 + r) * rank + c] : 0.0;
    }
    __syncthreads();

    double sumsq[kRowsB], amax[kRowsB], invq[kRowsB];
#pragma unroll
    for (int r = 0; r < kRowsB; ++r) {
        sumsq[r] = 0.0;
        amax[r] = 0.0;
        invq[r] = (WRITE && row0 + r < rows) ? inv_quantum[row0 + r] : 0.0;
    }

    for (int blk = b_begin + warp; blk < b_end; blk += kWarps) {
        const int64_t k0 = (int64_t)blk * 128 + lane;
        double z[kRowsB][4];
#pragma unroll
        for (int r = 0; r < kRowsB; ++r) {
            const bool rv = row0 + r < rows;
            const double* xr = X + (rv ? row0 + r : 0) * ldx;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + 32 * j;
                z[r][j] = (rv && k < n) ? __ldg(xr + k) : 0.0;
            }
        }
        for (int c0 = 0; c0 < rank4; c0 += 4) {
            double q[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t k = k0 + 32 * j;
                    q[c][j] = (c0 + c < rank && k < n) ? __ldg(Qt + (int64_t)(c0 + c) * ldq + k) : 0.0;
                }
#pragma unroll
            for (int r = 0; r < kRowsB; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double b = s_coef[r][c0 + c];
#pragma unroll
                    for (int j = 0; j < 4; ++j) z[r][j] = fma(-b, q[c][j], z[r][j]);
                }
        }
        bool flip[4];
        if (HAD) {
#pragma unroll
            for (int j = 0; j < 4; ++j) flip[j] = cell_flip(cell_offset + (uint64_t)(k0 + 32 * j));
        }
#pragma unroll
        for (int r = 0; r < kRowsB; ++r) {
            double a = z[r][0], b = z[r][1], c = z[r][2], d = z[r][3];
            sumsq[r] += a * a + b * b + c * c + d * d;
            if (HAD) {
                if (flip[0]) a = -a;
                if (flip[1]) b = -b;
                if (flip[2]) c = -c;
                if (flip[3]) d = -d;
                // element index e = 32 j + lane: bits 5,6 are in-thread, bits 0..4 across lanes
                double t0 = a + b, t1 = a - b, t2 = c + d, t3 = c - d;
                a = t0 + t2; b = t1 + t3; c = t0 - t2; d = t1 - t3;
#pragma unroll
                for (int m = 1; m < 32; m <<= 1) {
                    const bool up = lane & m;
                    double p;
                    p = __shfl_xor_sync(0xffffffffu, a, m); a = up ? p - a : a + p;
                    p = __shfl_xor_sync(0xffffffffu, b, m); b = up ? p - b : b + p;
                    p = __shfl_xor_sync(0xffffffffu, c, m); c = up ? p - c : c + p;
                    p = __shfl_xor_sync(0xffffffffu, d, m); d = up ? p - d : d + p;
                }
                a *= kHadScale; b *= kHadScale; c *= kHadScale; d *= kHadScale;
            }
            amax[r] = fmax(amax[r], fmax(fmax(fabs(a), fabs(b)), fmax(fabs(c), fabs(d))));
            if (WRITE) {
                const double zz[4] = {a, b, c, d};
                uint32_t word[NSR_MAX_SLICES] = {0, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double t = fmin(fmax(zz[j] * invq[r], -vmax), vmax);
                    int8_t dg[NSR_MAX_SLICES];
                    nsr_digits(__double2int_rn(t), n_slices, dg);
#pragma unroll
                    for (int s = 0; s < NSR_MAX_SLICES; ++s)
                        if (s < n_slices) word[s] |= (uint32_t)(uint8_t)dg[s] << (8 * j);
                }
                // lane l holds bytes of cells l+32j; regroup so lane l owns cells 4l..4l+3
                if (row0 + r < rows) {
#pragma unroll
                    for (int s = 0; s < NSR_MAX_SLICES; ++s) {
                        if (s < n_slices) {
                            uint32_t out = 0;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t w = __shfl_sync(0xffffffffu, word[s], (4 * lane + i) & 31);
                                out |= ((w >> (8 * (lane >> 3))) & 0xFFu) << (8 * i);
                            }
                            int8_t* dst = slices + ((int64_t)s * rows_alloc + row0 + r) * n_pad +
                                          (int64_t)blk * 128 + 4 * lane;
                            *reinterpret_cast<uint32_t*>(dst) = out;
                        }
                    }
                }
            }
        }
    }
    if (!WRITE) {
#pragma unroll
        for (int r = 0; r < kRowsB; ++r) {
            const double s = warp_sum(sumsq[r]);
            const double m = warp_max(amax[r]);
            if (lane == 0) { s_red[0][warp][r] = s; s_red[1][warp][r] = m; }
        }
        __syncthreads();
        if (threadIdx.x < kRowsB && row0 + threadIdx.x < rows) {
            double s = 0.0, m = 0.0;
            for (int w = 0; w < kWarps; ++w) {
                s += s_red[0][w][threadIdx.x];
                m = fmax(m, s_red[1][w][threadIdx.x]);
            }
            part_sumsq[(int64_t)ks * rows + row0 + threadIdx.x] = s;
            part_amax[(int64_t)ks * rows + row0 + threadIdx.x] = m;
        }
    }
}

__global__ void stats_finalize_kernel(const double* __restrict__ part_sumsq,
                                      const double* __restrict__ part_amax, int64_t rows,
                                      int ksplit, int64_t n, double vmax,
                                      double* __restrict__ var, double* __restrict__ quantum,
                                      double* __restrict__ inv_quantum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double s = 0.0, m = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) {
        s += part_sumsq[(int64_t)ks * rows + i];
        m = fmax(m, part_amax[(int64_t)ks * rows + i]);
    }
    double v = s / (double)n;
    if (v == 0.0) v = 1.0;                       // association.py:231,233
    var[i] = v;
    const double q = (m > 0.0 && isfinite(m)) ? m / vmax : 1.0;
    quantum[i] = q;
    inv_quantum[i] = 1.0 / q;
}

__global__ void unslice_kernel(const int8_t* __restrict__ slices, int64_t rows, int64_t rows_alloc,
                               int64_t n_pad, int n_slices, const double* __restrict__ quantum,
                               double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * n_pad) return;
    const int64_t r = i / n_pad, k = i % n_pad;
    int64_t v = 0;
    for (int s = 0; s < n_slices; ++s) v = v * 256 + slices[((int64_t)s * rows_alloc + r) * n_pad + k];
    out[i] = (double)v * quantum[r];
}

template <int CB>
void launch_coef(cudaStream_t st, dim3 grid, const double* X, int64_t rows, int64_t n, int64_t ldx,
                 const double* Qt, int rank, int64_t ldq, int c0, int nblk, int ksplit,
                 double* partial) {
    coef_kernel<CB><<<grid, kThreads, 0, st>>>(X, rows, n, ldx, Qt, rank, ldq, c0, nblk, ksplit,
                                               partial);
}

}  // namespace

int nsr_use_hadamard = 1;   // test hook (nsr_set_option)

extern "C" int64_t nsr_padded_cells(int64_t n) { return (n + NSR_KBLOCK - 1) / NSR_KBLOCK * NSR_KBLOCK; }

extern "C" int nsr_residualize(nsr_ctx* ctx, uintptr_t stream, const double* X, int64_t rows,
                               int64_t n, int64_t ldx, const double* Qt, int rank, int64_t ldq,
                               int n_slices, int8_t* slices, int64_t rows_alloc, int64_t n_pad,
                               double* quantum, double* var, double* coef) {
    NSR_REQUIRE(ctx != nullptr, "nsr_residualize: null context");
    NSR_REQUIRE(rows > 0 && n > 0 && ldx >= n, "nsr_residualize: bad shape rows=%lld n=%lld ldx=%lld",
                (long long)rows, (long long)n, (long long)ldx);
    NSR_REQUIRE(rank >= 0 && rank <= NSR_MAX_RANK, "nsr_residualize: rank %d outside [0,%d]", rank,
                NSR_MAX_RANK);
    NSR_REQUIRE(rank == 0 || (Qt != nullptr && ldq >= n), "nsr_residualize: bad covariate basis");
    NSR_REQUIRE(n_slices >= 2 && n_slices <= NSR_MAX_SLICES, "nsr_residualize: n_slices %d", n_slices);
    NSR_REQUIRE(n_pad == nsr_padded_cells(n) && rows_alloc >= rows,
                "nsr_residualize: n_pad/rows_alloc inconsistent");
    NSR_REQUIRE(((uintptr_t)slices & 15) == 0, "nsr_residualize: slices must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    NSR_CHECK(cudaSetDevice(ctx->device));

    const int nblk = (int)(n_pad / 128);
    auto pick_split = [&](int64_t groups) {
        int64_t ks = (2 * (int64_t)ctx->sm_count + groups - 1) / groups;
        if (ks < 1) ks = 1;
        if (ks > nblk) ks = nblk;
        if (ks > 64) ks = 64;
        return (int)ks;
    };
    const int64_t groups_a = (rows + kRowsA - 1) / kRowsA;
    const int64_t groups_b = (rows + kRowsB - 1) / kRowsB;
    const int ks_a = pick_split(groups_a), ks_b = pick_split(groups_b);

    // scratch: coef partials | sumsq partials | amax partials | inv_quantum | coef (if caller passed none)
    const size_t n_part = (size_t)ks_a * rows * (rank > 0 ? rank : 1);
    const size_t n_coef = (size_t)rows * (rank > 0 ? rank : 1);
    const size_t total = n_part + 2 * (size_t)ks_b * rows + rows + n_coef;
    void* scratch = nullptr;
    if (nsr_scratch(ctx, total * sizeof(double), &scratch)) return 1;
    double* partial = (double*)scratch;
    double* p_sumsq = partial + n_part;
    double* p_amax = p_sumsq + (size_t)ks_b * rows;
    double* invq = p_amax + (size_t)ks_b * rows;
    double* coef_buf = coef ? coef : invq + rows;

    if (rank > 0) {
        const dim3 grid((unsigned)groups_a, (unsigned)ks_a);
        for (int c0 = 0; c0 < rank; c0 += 12) {
            const int left = rank - c0;
            if (left <= 4) launch_coef<4>(st, grid, X, rows, n, ldx, Qt, rank, ldq, c0, nblk, ks_a, partial);
            else if (left <= 8) launch_coef<8>(st, grid, X, rows, n, ldx, Qt, rank, ldq, c0, nblk, ks_a, partial);
            else launch_coef<12>(st, grid, X, rows, n, ldx, Qt, rank, ldq, c0, nblk, ks_a, partial);
        }
        const int64_t count = rows * rank;
        coef_reduce_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(partial, count, ks_a, coef_buf);
    }
    const double vmax = nsr_vmax(n_slices);
    const dim3 gridb((unsigned)groups_b, (unsigned)ks_b);
    if (nsr_use_hadamard)
        residual_kernel<false, true><<<gridb, kThreads, 0, st>>>(X, rows, n, ldx, Qt, rank, ldq, coef_buf, nblk, ks_b, 0, p_sumsq, p_amax, nullptr, n_slices, vmax, nullptr, rows_alloc, n_pad);
    else
        residual_kernel<false, false><<<gridb, kThreads, 0, st>>>(X, rows, n, ldx, Qt, rank, ldq, coef_buf, nblk, ks_b, 0, p_sumsq, p_amax, nullptr, n_slices, vmax, nullptr, rows_alloc, n_pad);
    stats_finalize_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(p_sumsq, p_amax, rows, ks_b, n, vmax, var, quantum, invq);
    if (nsr_use_hadamard)
        residual_kernel<true, true><<<gridb, kThreads, 0, st>>>(X, rows, n, ldx, Qt, rank, ldq, coef_buf, nblk, ks_b, 0, nullptr, nullptr, invq, n_slices, vmax, slices, rows_alloc, n_pad);
    else
        residual_kernel<true, false><<<gridb, kThreads, 0, st>>>(X, rows, n, ldx, Qt, rank, ldq, coef_buf, nblk, ks_b, 0, nullptr, nullptr, invq, n_slices, vmax, slices, rows_alloc, n_pad);
    NSR_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int nsr_unslice(nsr_ctx* ctx, uintptr_t stream, const int8_t* slices, int64_t rows,
                           int64_t rows_alloc, int64_t n_pad, int n_slices, const double* quantum,
                           double* out) {
    NSR_REQUIRE(ctx != nullptr, "nsr_unslice: null context");
    NSR_CHECK(cudaSetDevice(ctx->device));
    const int64_t count = rows * n_pad;
    unslice_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        slices, rows, rows_alloc, n_pad, n_slices, quantum, out);
    NSR_CHECK(cudaGetLastError());
    return 0;
}
