// Covariate projection + quantisation ("residualize"): the HBM-streaming half of the path.
//
// Reference: association.py:226-233
//     ccx = dci @ (dc @ dx.T);  dx1 = dx - ccx.T @ dc;  var = mean(dx1**2), 0 -> 1
// Here the host supplies Qt, an orthonormal basis (rank x n) of the row space of dc, so
//     coef = X Qt^T (pass A),   z = X - coef Qt (pass B).
// Pass B also applies a sign-randomised orthonormal 128-point Walsh-Hadamard transform along
// cells (z' = z D H: inner products over cells are unchanged, rows become near-Gaussian so a
// fixed-point row scale wastes no bits on outliers) and writes round(z'/quantum) as balanced
// base-256 int8 digit planes - the operand format of the tensor-core contraction.
//
// Two passes over X (algorithmic traffic 8 B read + S B written per element, actual 16 + S):
//   A  coef and sum(x^2) per row -> rms estimate sqrt((sum x^2 - |coef|^2)/n) (Qt orthonormal)
//   B  z, z', exact var = mean(z^2), exact max|z'|, digits with quantum = kKappa*rms_est/vmax
// A row whose max|z'| exceeds kKappa*rms_est (probability ~2e-9 per element for Gaussianised
// rows, certain for degenerate ones) is re-quantised by a third, sparse pass with
// quantum = max|z'|/vmax.
//
// Layout: all 8 warps of a CTA walk the same 128-cell blocks (so the covariate block is an L1
// hit for 7 of them) and own different rows.  Lane l holds cells l, l+32, l+64, l+96 of the
// block: every global load is a fully coalesced 256 B row segment.  For the transform each warp
// transposes its 8 rows x 128 cells through shared memory so that a lane owns 32 consecutive
// cells of one row: 5 butterfly stages in registers, 2 by shuffle, and 32-byte digit stores.
#include "nsr_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsA = 4;                  // rows per warp, pass A
constexpr int kRowsB = 8;                  // rows per warp, pass B
constexpr int kSegPitch = 33;              // doubles per 32-cell segment in smem (+1: bank spread)
constexpr double kHadScale = 0.088388347648318440550;   // 1/sqrt(128)
constexpr double kKappa = 6.0;

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

// ---- pass A ------------------------------------------------------------------------------
// partial[ks][row][c] = sum over the CTA's cells of X[row][k] Qt[c0+c][k];  psq[ks][row] = sum x^2
template <int CB, bool FULL>
__device__ __forceinline__ void coef_block(const double* const (&xp)[kRowsA], const double* qp, int64_t ldq, int nq,
                                           int64_t k0, int64_t n, bool first_chunk, double (&acc)[kRowsA][CB],
                                           double (&sq)[kRowsA]) {
    double x[kRowsA][4];
#pragma unroll
    for (int r = 0; r < kRowsA; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[r][j] = (FULL || k0 + 32 * j < n) ? __ldg(xp[r] + k0 + 32 * j) : 0.0;
    const double* qc = qp + k0;
#pragma unroll
    for (int c = 0; c < CB; ++c) {
        double q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = (c < nq && (FULL || k0 + 32 * j < n)) ? __ldg(qc + 32 * j) : 0.0;
        qc += ldq;
#pragma unroll
        for (int r = 0; r < kRowsA; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][c] = fma(x[r][j], q[j], acc[r][c]);
    }
    if (first_chunk) {
#pragma unroll
        for (int r = 0; r < kRowsA; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) sq[r] = fma(x[r][j], x[r][j], sq[r]);
    }
}

template <int CB>
__global__ void __launch_bounds__(kThreads, 2)
coef_kernel(const double* __restrict__ X, int64_t rows, int64_t n, int64_t ldx,
            const double* __restrict__ Qt, int rank, int64_t ldq, int c0, int nblk, int ksplit,
            double* __restrict__ partial, double* __restrict__ psq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = ((int64_t)blockIdx.x * kWarps + warp) * kRowsA;
    const int ks = blockIdx.y;
    const int b_begin = (int)((int64_t)nblk * ks / ksplit);
    const int b_end = (int)((int64_t)nblk * (ks + 1) / ksplit);
    if (row0 >= rows) return;
    const int nblk_full = (int)(n / 128);
    const int nq = rank - c0 < CB ? rank - c0 : CB;

    const double* xp[kRowsA];
#pragma unroll
    for (int r = 0; r < kRowsA; ++r) xp[r] = X + (row0 + r < rows ? row0 + r : row0) * ldx + lane;
    const double* qp = Qt + (int64_t)c0 * ldq + lane;
    double acc[kRowsA][CB], sq[kRowsA];
#pragma unroll
    for (int r = 0; r < kRowsA; ++r) {
        sq[r] = 0.0;
#pragma unroll
        for (int c = 0; c < CB; ++c) acc[r][c] = 0.0;
    }
    int blk = b_begin;
    const int fast_end = b_end < nblk_full ? b_end : nblk_full;
    for (; blk < fast_end; ++blk) coef_block<CB, true>(xp, qp, ldq, nq, (int64_t)blk * 128, n, c0 == 0, acc, sq);
    for (; blk < b_end; ++blk) coef_block<CB, false>(xp, qp, ldq, nq, (int64_t)blk * 128, n - lane, c0 == 0, acc, sq);
#pragma unroll
    for (int r = 0; r < kRowsA; ++r) {
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            const double v = warp_sum(acc[r][c]);
            if (lane == 0 && row0 + r < rows && c < nq)
                partial[((int64_t)ks * rows + row0 + r) * rank + c0 + c] = v;
        }
        if (c0 == 0) {
            const double v = warp_sum(sq[r]);
            if (lane == 0 && row0 + r < rows) psq[(int64_t)ks * rows + row0 + r] = v;
        }
    }
}

// rank == 0: only sum x^2 is needed
__global__ void __launch_bounds__(kThreads, 2)
sumsq_kernel(const double* __restrict__ X, int64_t rows, int64_t n, int64_t ldx, int nblk, int ksplit,
             double* __restrict__ psq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarps + warp;
    const int ks = blockIdx.y;
    if (row >= rows) return;
    const int64_t k_begin = (int64_t)nblk * ks / ksplit * 128, k_end = (int64_t)nblk * (ks + 1) / ksplit * 128;
    double s = 0.0;
    for (int64_t k = k_begin + lane; k < k_end && k < n; k += 32) {
        const double x = __ldg(X + row * ldx + k);
        s = fma(x, x, s);
    }
    s = warp_sum(s);
    if (lane == 0) psq[(int64_t)ks * rows + row] = s;
}

// coef[row][c] = sum_ks partial; inv_quantum estimate from rms_est = sqrt((sum x^2 - |coef|^2)/n)
__global__ void coef_finalize_kernel(const double* __restrict__ partial, const double* __restrict__ psq,
                                     int64_t rows, int rank, int ksplit, int64_t n, double vmax,
                                     double* __restrict__ coef, double* __restrict__ invq_est) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    double c2 = 0.0;
    for (int c = 0; c < rank; ++c) {
        double s = 0.0;
        for (int ks = 0; ks < ksplit; ++ks) s += partial[((int64_t)ks * rows + row) * rank + c];   // fixed order
        coef[row * rank + c] = s;
        c2 = fma(s, s, c2);
    }
    double sq = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) sq += psq[(int64_t)ks * rows + row];
    const double ms = (sq - c2) / (double)n;
    const double est = (ms > 0.0 && isfinite(ms)) ? kKappa * sqrt(ms) : 0.0;
    invq_est[row] = est > 0.0 ? vmax / est : 0.0;          // 0 -> digits 0, row goes to the fix-up pass
}

// ---- pass B ------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void store_digits(const double (&v)[32], double invq, double vmax,
                                             int8_t* __restrict__ dst, int64_t plane_stride) {
    uint32_t w[S][8];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int i = 0; i < 8; ++i) w[s][i] = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const double t = fmin(fmax(v[i] * invq, -vmax), vmax);
        int32_t q = __double2int_rn(t);
#pragma unroll
        for (int s = S - 1; s >= 1; --s) {
            // balanced low digit d = sext8(q & 0xFF); (q - d) >> 8 == (q + 128) >> 8
            w[s][i >> 2] |= (uint32_t)(q & 0xFF) << (8 * (i & 3));
            q = (q + 128) >> 8;
        }
        w[0][i >> 2] |= (uint32_t)(q & 0xFF) << (8 * (i & 3));
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        uint4* p = reinterpret_cast<uint4*>(dst + (int64_t)s * plane_stride);
        p[0] = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
        p[1] = make_uint4(w[s][4], w[s][5], w[s][6], w[s][7]);
    }
}

// row_list == nullptr: logical row == row.  Otherwise the kernel handles rows row_list[0..*row_count).
template <int S, bool HAD>
__global__ void __launch_bounds__(kThreads, 2)
residual_kernel(const double* __restrict__ X, int64_t rows, int64_t n, int64_t ldx,
                const double* __restrict__ Qt, int rank, int64_t ldq, const double* __restrict__ coef,
                const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_count,
                int nblk, int ksplit, uint64_t cell_offset, const double* __restrict__ inv_quantum,
                double vmax, double* __restrict__ part_sumsq, double* __restrict__ part_amax,
                int8_t* __restrict__ slices, int64_t rows_alloc, int64_t n_pad) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_logical = row_list ? (int64_t)*row_count : rows;
    const int64_t l0 = ((int64_t)blockIdx.x * kWarps + warp) * kRowsB;    // first logical row of this warp
    const int ks = blockIdx.y;
    const int b_begin = (int)((int64_t)nblk * ks / ksplit);
    const int b_end = (int)((int64_t)nblk * (ks + 1) / ksplit);
    if ((int64_t)blockIdx.x * kWarps * kRowsB >= n_logical) return;        // whole CTA idle
    const int rank4 = (rank + 3) & ~3;

    double* s_coef = smem + warp * (kRowsB * NSR_MAX_RANK + 32 * kSegPitch);   // [kRowsB][NSR_MAX_RANK]
    double* s_z = s_coef + kRowsB * NSR_MAX_RANK;                               // [32 segments][kSegPitch]
    int64_t row_of[kRowsB];
#pragma unroll
    for (int r = 0; r < kRowsB; ++r) {
        const int64_t l = l0 + r;
        row_of[r] = (l < n_logical) ? (row_list ? (int64_t)row_list[l] : l) : -1;
    }
    for (int i = lane; i < kRowsB * NSR_MAX_RANK; i += 32) {
        const int r = i / NSR_MAX_RANK, c = i % NSR_MAX_RANK;
        const int64_t l = l0 + r;
        const int64_t row = (l < n_logical) ? (row_list ? (int64_t)row_list[l] : l) : -1;
        s_coef[i] = (row >= 0 && c < rank) ? coef[row * rank + c] : 0.0;
    }
    __syncwarp();

    // phase-2 ownership: lane -> (row slot lane>>2, cells 32*(lane&3) .. +32 of the block)
    const int my_slot = lane >> 2, my_quarter = lane & 3;
    int64_t my_row = -1;
#pragma unroll
    for (int r = 0; r < kRowsB; ++r)
        if (r == my_slot) my_row = row_of[r];
    const double my_invq = (my_row >= 0 && inv_quantum) ? inv_quantum[my_row] : 0.0;
    double my_amax = 0.0;
    double sumsq[kRowsB];
#pragma unroll
    for (int r = 0; r < kRowsB; ++r) sumsq[r] = 0.0;

    const int nblk_full = (int)(n / 128);
    const double* xp[kRowsB];
#pragma unroll
    for (int r = 0; r < kRowsB; ++r) xp[r] = X + (row_of[r] >= 0 ? row_of[r] : 0) * ldx + lane;
    const double* qbase = Qt + lane;
    for (int blk = b_begin; blk < b_end; ++blk) {
        const int64_t k0 = (int64_t)blk * 128;
        const bool full = blk < nblk_full;
        const int64_t n_lane = n - lane;                 // cell (k0 + 32 j + lane) < n  <=>  k0 + 32 j < n_lane
        bool flip[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) flip[j] = HAD && cell_flip(cell_offset + (uint64_t)(k0 + lane + 32 * j));
        // ---- phase 1: residuals of 8 rows (two halves of 4 to bound registers) -> smem
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            double z[4][4];
            if (full) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j) z[r][j] = __ldg(xp[4 * half + r] + k0 + 32 * j);
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        z[r][j] = (k0 + 32 * j < n_lane) ? __ldg(xp[4 * half + r] + k0 + 32 * j) : 0.0;
            }
            const double* qc = qbase + k0;
            for (int c0 = 0; c0 < rank4; c0 += 4) {
                double q[4][4];
                if (full && c0 + 4 <= rank) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 4; ++j) q[c][j] = __ldg(qc + (int64_t)c * ldq + 32 * j);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            q[c][j] = (c0 + c < rank && k0 + 32 * j < n_lane) ? __ldg(qc + (int64_t)c * ldq + 32 * j) : 0.0;
                }
                qc += 4 * ldq;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double b = s_coef[(4 * half + r) * NSR_MAX_RANK + c0 + c];
#pragma unroll
                        for (int j = 0; j < 4; ++j) z[r][j] = fma(-b, q[c][j], z[r][j]);
                    }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double v = z[r][j];
                    sumsq[4 * half + r] = fma(v, v, sumsq[4 * half + r]);
                    s_z[(4 * (4 * half + r) + j) * kSegPitch + lane] = flip[j] ? -v : v;
                }
        }
        __syncwarp();
        // ---- phase 2: 32 consecutive cells of one row per lane
        double v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = s_z[lane * kSegPitch + i];
        __syncwarp();
        if (HAD) {
#pragma unroll
            for (int h = 1; h < 32; h <<= 1)
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if ((i & h) == 0) {
                        const double a = v[i], b = v[i + h];
                        v[i] = a + b;
                        v[i + h] = a - b;
                    }
#pragma unroll
            for (int m = 1; m <= 2; m <<= 1) {
                const bool up = lane & m;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const double p = __shfl_xor_sync(0xffffffffu, v[i], m);
                    v[i] = up ? p - v[i] : v[i] + p;
                }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= kHadScale;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) my_amax = fmax(my_amax, fabs(v[i]));
        if (slices != nullptr && my_row >= 0)
            store_digits<S>(v, my_invq, vmax,
                            slices + my_row * n_pad + (int64_t)blk * 128 + 32 * my_quarter,
                            rows_alloc * n_pad);
    }
    if (part_sumsq != nullptr) {
#pragma unroll
        for (int r = 0; r < kRowsB; ++r) {
            const double s = warp_sum(sumsq[r]);
            if (lane == 0 && row_of[r] >= 0) part_sumsq[(int64_t)ks * rows + row_of[r]] = s;
        }
        my_amax = fmax(my_amax, __shfl_xor_sync(0xffffffffu, my_amax, 1));
        my_amax = fmax(my_amax, __shfl_xor_sync(0xffffffffu, my_amax, 2));
        if (my_quarter == 0 && my_row >= 0) part_amax[(int64_t)ks * rows + my_row] = my_amax;
    }
}

// var, final quantum, and the list of rows whose digits overflowed the estimated scale
__global__ void stats_finalize_kernel(const double* __restrict__ part_sumsq,
                                      const double* __restrict__ part_amax, int64_t rows,
                                      int ksplit, int64_t n, double vmax,
                                      double* __restrict__ var, double* __restrict__ quantum,
                                      double* __restrict__ inv_quantum, int32_t* __restrict__ fix_list,
                                      int32_t* __restrict__ fix_count) {
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
    const double iq = inv_quantum[i];
    if (m > 0.0 && isfinite(m) && (iq == 0.0 || !(m * iq <= vmax))) {     // iq == 0: no usable estimate
        const double q = m / vmax;
        quantum[i] = q;
        inv_quantum[i] = 1.0 / q;
        fix_list[atomicAdd(fix_count, 1)] = (int32_t)i;
    } else {
        quantum[i] = iq > 0.0 ? 1.0 / iq : 1.0;
    }
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

constexpr int kSmemB = kWarps * (kRowsB * NSR_MAX_RANK + 32 * kSegPitch) * (int)sizeof(double);

template <int S, bool HAD>
int launch_residual(cudaStream_t st, dim3 grid, const double* X, int64_t rows, int64_t n, int64_t ldx,
                    const double* Qt, int rank, int64_t ldq, const double* coef, const int32_t* row_list,
                    const int32_t* row_count, int nblk, int ksplit, const double* invq, double vmax,
                    double* p_sumsq, double* p_amax, int8_t* slices, int64_t rows_alloc, int64_t n_pad) {
    auto kern = residual_kernel<S, HAD>;
    static bool attr_done = false;
    if (!attr_done) {
        NSR_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemB));
        attr_done = true;
    }
    kern<<<grid, kThreads, kSmemB, st>>>(X, rows, n, ldx, Qt, rank, ldq, coef, row_list, row_count, nblk, ksplit,
                                         0, invq, vmax, p_sumsq, p_amax, slices, rows_alloc, n_pad);
    return 0;
}

}  // namespace

int nsr_use_hadamard = 1;   // test hook (nsr_set_option)

extern "C" int64_t nsr_padded_cells(int64_t n) { return (n + NSR_KBLOCK - 1) / NSR_KBLOCK * NSR_KBLOCK; }

extern "C" int nsr_residualize(nsr_ctx* ctx, uintptr_t stream, const double* X, int64_t rows,
                               int64_t n, int64_t ldx, const double* Qt, int rank, int64_t ldq,
                               int n_slices, int8_t* slices, int64_t rows_alloc, int64_t n_pad,
                               double* quantum, double* var, double* coef) {
    NSR_REQUIRE(ctx != nullptr, "nsr_residualize: null context");
    NSR_REQUIRE(rows > 0 && rows < (1ll << 31) && n > 0 && ldx >= n,
                "nsr_residualize: bad shape rows=%lld n=%lld ldx=%lld", (long long)rows, (long long)n,
                (long long)ldx);
    NSR_REQUIRE(rank >= 0 && rank <= NSR_MAX_RANK, "nsr_residualize: rank %d outside [0,%d]", rank,
                NSR_MAX_RANK);
    NSR_REQUIRE(rank == 0 || (Qt != nullptr && ldq >= n), "nsr_residualize: bad covariate basis");
    NSR_REQUIRE(n_slices == 3 || n_slices == 4, "nsr_residualize: n_slices %d (3 or 4)", n_slices);
    NSR_REQUIRE(n_pad == nsr_padded_cells(n) && rows_alloc >= rows,
                "nsr_residualize: n_pad/rows_alloc inconsistent");
    NSR_REQUIRE(((uintptr_t)slices & 15) == 0, "nsr_residualize: slices must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    NSR_CHECK(cudaSetDevice(ctx->device));

    const int nblk = (int)(n_pad / 128);
    auto pick_split = [&](int64_t groups) {
        int64_t ks = (4 * (int64_t)ctx->sm_count + groups - 1) / groups;
        if (ks < 1) ks = 1;
        if (ks > nblk) ks = nblk;
        if (ks > 64) ks = 64;
        return (int)ks;
    };
    const int64_t groups_a = (rows + kWarps * kRowsA - 1) / (kWarps * kRowsA);
    const int64_t groups_s = (rows + kWarps - 1) / kWarps;
    const int64_t groups_b = (rows + kWarps * kRowsB - 1) / (kWarps * kRowsB);
    const int ks_a = pick_split(rank ? groups_a : groups_s), ks_b = pick_split(groups_b);
    const int rk = rank > 0 ? rank : 1;

    // scratch (doubles): coef partials | sum-x^2 partials | sumsq partials | amax partials | inv_quantum
    //                    | coef (if the caller passed none) ; then int32: fix_count(+pad) | fix_list
    const size_t n_part = (size_t)ks_a * rows * rk, n_psq = (size_t)ks_a * rows;
    const size_t n_stat = (size_t)ks_b * rows, n_coef = (size_t)rows * rk;
    const size_t n_dbl = n_part + n_psq + 2 * n_stat + rows + n_coef;
    void* scratch = nullptr;
    if (nsr_scratch(ctx, n_dbl * sizeof(double) + (size_t)(rows + 4) * sizeof(int32_t), &scratch)) return 1;
    double* partial = (double*)scratch;
    double* psq = partial + n_part;
    double* p_sumsq = psq + n_psq;
    double* p_amax = p_sumsq + n_stat;
    double* invq = p_amax + n_stat;
    double* coef_buf = coef ? coef : invq + rows;
    int32_t* fix_count = (int32_t*)((double*)scratch + n_dbl);
    int32_t* fix_list = fix_count + 4;
    const double vmax = nsr_vmax(n_slices);

    NSR_CHECK(cudaMemsetAsync(fix_count, 0, sizeof(int32_t), st));
    if (rank > 0) {
        const dim3 grid((unsigned)groups_a, (unsigned)ks_a);
        for (int c0 = 0; c0 < rank; c0 += 8) {
            if (rank - c0 <= 4)
                coef_kernel<4><<<grid, kThreads, 0, st>>>(X, rows, n, ldx, Qt, rank, ldq, c0, nblk, ks_a, partial, psq);
            else
                coef_kernel<8><<<grid, kThreads, 0, st>>>(X, rows, n, ldx, Qt, rank, ldq, c0, nblk, ks_a, partial, psq);
        }
    } else {
        sumsq_kernel<<<dim3((unsigned)groups_s, (unsigned)ks_a), kThreads, 0, st>>>(X, rows, n, ldx, nblk, ks_a, psq);
    }
    coef_finalize_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(partial, psq, rows, rank, ks_a, n, vmax,
                                                                       coef_buf, invq);
    const dim3 gridb((unsigned)groups_b, (unsigned)ks_b);
    const bool had = nsr_use_hadamard != 0;
    int rc;
#define NSR_LAUNCH_B(GRID, LIST, COUNT, KS, PS, PA)                                                          \
    (n_slices == 3 ? (had ? launch_residual<3, true>(st, GRID, X, rows, n, ldx, Qt, rank, ldq, coef_buf, LIST, COUNT, nblk, KS, invq, vmax, PS, PA, slices, rows_alloc, n_pad)   \
                          : launch_residual<3, false>(st, GRID, X, rows, n, ldx, Qt, rank, ldq, coef_buf, LIST, COUNT, nblk, KS, invq, vmax, PS, PA, slices, rows_alloc, n_pad)) \
                   : (had ? launch_residual<4, true>(st, GRID, X, rows, n, ldx, Qt, rank, ldq, coef_buf, LIST, COUNT, nblk, KS, invq, vmax, PS, PA, slices, rows_alloc, n_pad)   \
                          : launch_residual<4, false>(st, GRID, X, rows, n, ldx, Qt, rank, ldq, coef_buf, LIST, COUNT, nblk, KS, invq, vmax, PS, PA, slices, rows_alloc, n_pad)))
    rc = NSR_LAUNCH_B(gridb, nullptr, nullptr, ks_b, p_sumsq, p_amax);
    if (rc) return rc;
    stats_finalize_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(p_sumsq, p_amax, rows, ks_b, n, vmax, var,
                                                                        quantum, invq, fix_list, fix_count);
    // sparse fix-up: CTAs beyond the (device-side) count exit at once
    rc = NSR_LAUNCH_B(gridb, fix_list, fix_count, ks_b, nullptr, nullptr);
#undef NSR_LAUNCH_B
    if (rc) return rc;
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
