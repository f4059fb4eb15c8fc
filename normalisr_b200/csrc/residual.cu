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
// Both passes put the skinny float64 products on the FP64 tensor cores (mma.sync m8n8k4, "DMMA"):
// one warp instruction does 256 FMAs, so the kernels issue ~0.1 (A) / ~30 (B, mostly the
// butterflies and the digit extraction) instructions per element instead of ~60-110 with scalar
// FMAs, which left them issue-bound at a quarter of the HBM rate (profiles/r01c_project_ncu.md).
//   fragment layouts (g = lane/4, t = lane%4):  A[8x4]: (row g, k t)   B[4x8]: (k t, col g)
//                                               C[8x8]: (row g, cols 2t, 2t+1)
//   pass A: M = 8 rows, N = 8 covariates, K = cells.  Lane (g,t) loads 4 consecutive cells of row g
//           (a quad reads a full 128 B line) and uses them as the K index of 4 successive MMAs.
//   pass B: M = 8 rows, N = 8 cells, K = 4 covariates: C = X tile, A = -coef, B = Qt tile, so the
//           residual comes out in the accumulator layout: after 16 tiles lane (g,t) owns the 32
//           cells {8u + 2t + e} of row g.  5 butterfly stages are then in registers, 2 by shuffle.
// Cells are stored in that (fixed, row-independent) order: position 32t + 2u + e of the 128-cell
// block holds transformed index 8u + 2t + e.  The contraction sums over all positions, so any
// fixed permutation is as good as the natural order, and every lane stores 32 contiguous bytes.
#include "nsr_common.cuh"

extern int nsr_prefetch;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsW = 8;                  // rows per warp (one MMA row tile)
constexpr double kHadScale = 0.088388347648318440550;   // 1/sqrt(128)
constexpr double kKappa = 6.0;
constexpr int kQPitch = 132;
constexpr int kXPitch = 136;                // row pitch (doubles) of a warp's staged X tile
constexpr int kResidualSmem = (2 * 16 * kQPitch + kWarps * kRowsW * kXPitch) * (int)sizeof(double);   // 103,424 B: 2 CTAs / SM

// Sign pattern D of the mix z' = z D H: cell m of every 128-cell block (m = 8u + 2t + e in the accumulator
// layout, t = lane % 4) is negated when bit 2u + e of kSignMask is set.  The pattern is the same for every
// block and lane, so inside the kernel it is a compile-time property of the register index: the negations
// fold into the operand modifiers of the first butterfly stage and cost nothing.  (Any fixed diagonal of
// +-1 keeps all inner products; it only has to break alignments between the data and the Walsh functions.)
constexpr uint32_t kSignMask = 0x9e3779b9u;
__host__ __device__ constexpr bool reg_flip(int i) { return (kSignMask >> i) & 1u; }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
// D(8x8) += A(8x4) B(4x8), float64, one warp
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <bool VEC>
__device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
    if (VEC) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(p));
        const double2 b = __ldg(reinterpret_cast<const double2*>(p + 2));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = __ldg(p + e);
    }
}
__device__ __forceinline__ void load4_tail(const double* p, int64_t k, int64_t n, double (&v)[4]) {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (k + e < n) ? __ldg(p + e) : 0.0;
}

// ---- pass A ------------------------------------------------------------------------------
// partial[ks][row][c0 .. c0+8*NQ) = sum over this CTA's cells of X[row][k] Qt[c][k];  psq = sum x^2
template <int NQ, bool VEC>
__global__ void __launch_bounds__(kThreads, 3)
coef_mma_kernel(const double* __restrict__ X, int64_t rows, int64_t n, int64_t ldx,
                const double* __restrict__ Qt, int rank, int64_t ldq, int c0, int ksplit,
                double* __restrict__ partial, double* __restrict__ psq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t row_w = ((int64_t)blockIdx.x * kWarps + warp) * kRowsW;
    if (row_w >= rows) return;
    const int64_t row = row_w + g;
    const bool valid = row < rows;
    const double* xr = X + (valid ? row : row_w) * ldx + 4 * t;
    const double* qr[NQ];
    bool qok[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int c = c0 + 8 * q + g;
        qok[q] = c < rank;
        qr[q] = Qt + (int64_t)(qok[q] ? c : 0) * ldq + 4 * t;
    }
    const int64_t n16 = (n + 15) / 16, n16_full = n / 16;
    const int64_t kb = n16 * blockIdx.y / ksplit, ke = n16 * (blockIdx.y + 1) / ksplit;
    const int64_t ke_fast = ke < n16_full ? ke : n16_full;

    double acc[NQ][2], sq = 0.0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q][0] = acc[q][1] = 0.0;
    int64_t kk = kb;
#pragma unroll 4
    for (; kk < ke_fast; ++kk) {
        double xv[4], qv[NQ][4];
        load4<VEC>(xr + kk * 16, xv);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (qok[q]) load4<VEC>(qr[q] + kk * 16, qv[q]);
            else qv[q][0] = qv[q][1] = qv[q][2] = qv[q][3] = 0.0;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) dmma(acc[q][0], acc[q][1], xv[e], qv[q][e]);
            sq = fma(xv[e], xv[e], sq);
        }
    }
    for (; kk < ke; ++kk) {                                // ragged last group of 16 cells
        double xv[4], qv[NQ][4];
        const int64_t k = kk * 16 + 4 * t;
        load4_tail(xr + kk * 16, k, n, xv);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (qok[q]) load4_tail(qr[q] + kk * 16, k, n, qv[q]);
            else qv[q][0] = qv[q][1] = qv[q][2] = qv[q][3] = 0.0;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) dmma(acc[q][0], acc[q][1], xv[e], qv[q][e]);
            sq = fma(xv[e], xv[e], sq);
        }
    }
    // accumulator: row g, covariates c0 + 8q + 2t, +1
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = c0 + 8 * q + 2 * t + e;
            if (valid && c < rank) partial[((int64_t)blockIdx.y * rows + row) * rank + c] = acc[q][e];
        }
    if (c0 == 0) {
        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        if (t == 0 && valid) psq[(int64_t)blockIdx.y * rows + row] = sq;
    }
}

// rank == 0: only sum x^2 is needed
__global__ void __launch_bounds__(kThreads, 2)
sumsq_kernel(const double* __restrict__ X, int64_t rows, int64_t n, int64_t ldx, int ksplit,
             double* __restrict__ psq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarps + warp;
    const int ks = blockIdx.y;
    if (row >= rows) return;
    const int64_t nblk = (n + 127) / 128;
    const int64_t k_begin = nblk * ks / ksplit * 128, k_end = nblk * (ks + 1) / ksplit * 128;
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
                                     double* __restrict__ coef, double* __restrict__ invq_est,
                                     int* __restrict__ exact_status) {
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
    // exact single-plane path: the raw row stands in for its residual in the cross products, which
    // amplifies the partner's quantisation error by sqrt(sum x^2 / sum res^2): refuse beyond 2x
    if (exact_status != nullptr && !(sq <= 4.0 * (sq - c2))) atomicOr(exact_status, 2);
    const double est = (ms > 0.0 && isfinite(ms)) ? kKappa * sqrt(ms) : 0.0;
    invq_est[row] = est > 0.0 ? vmax / est : 0.0;          // 0 -> digits 0, row goes to the fix-up pass
}

// ---- pass B ------------------------------------------------------------------------------
// Quantise 32 values and write their balanced base-256 digits.  q = round(v * invq) comes from one fused
// multiply-add with the "magic" constant 1.5 * 2^52: the sum's low mantissa word IS the integer (round
// to nearest even, the same rounding as cvt.rni), and adding 128 per digit in the same constant makes
// every byte of the word a digit + 128 with no borrows between bytes (sum_k d_k 256^k + sum_k 128 256^k),
// so the digits are byte transposes and one XOR 0x80 per word instead of shifts and sign extensions.
// A value beyond the representable range wraps (the old cvt saturated): such rows are flagged from the
// exact maximum and rewritten by the fix-up pass either way.
template <int S>
__device__ __forceinline__ void store_digits(const double (&v)[32], double invq, int8_t* __restrict__ dst,
                                             int64_t plane_stride, uint32_t (&energy)[S]) {
    constexpr uint32_t kBias = S == 1 ? 0x80u : (S == 2 ? 0x8080u : (S == 3 ? 0x808080u : 0x80808080u));
    const double magic = 6755399441055744.0 + (double)kBias;
    uint32_t w[S][8];
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
        uint32_t q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = (uint32_t)__double2loint(fma(v[4 * i4 + j], invq, magic));
        // byte k of q[0..3] -> one word: two interleaves, then one pick per plane
        const uint32_t lo01 = __byte_perm(q[0], q[1], 0x5140), lo23 = __byte_perm(q[2], q[3], 0x5140);   // bytes 0, 1
        const uint32_t hi01 = __byte_perm(q[0], q[1], 0x7362), hi23 = __byte_perm(q[2], q[3], 0x7362);   // bytes 2, 3
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int k = S - 1 - s;                        // plane 0 is the most significant digit
            const uint32_t a = k < 2 ? lo01 : hi01, b = k < 2 ? lo23 : hi23;
            w[s][i4] = __byte_perm(a, b, (k & 1) ? 0x7632 : 0x5410) ^ 0x80808080u;
        }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {
        uint4* p = reinterpret_cast<uint4*>(dst + (int64_t)s * plane_stride);
        p[0] = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
        p[1] = make_uint4(w[s][4], w[s][5], w[s][6], w[s][7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) energy[s] = (uint32_t)__dp4a((int)w[s][i], (int)w[s][i], (int)energy[s]);   // sum d^2
    }
}

// row_list == nullptr: logical row == row.  Otherwise the kernel handles rows row_list[0..*row_count).
// NCH = covariate chunks of 4 handled from shared memory (rank <= 16 -> ceil(rank/4); larger ranks
// use NCH = 4 plus the global-memory groups).
// RAW (S = 1, nsr_residualize_exact): the plane holds the Hadamard mix of the RAW row, which for rows
// of small integers (binary groupings) is itself a small integer - stored exactly, quantum 1/sqrt(128);
// the residual is still formed for the exact variance.  *exact_status |= 1 if a value is not an integer
// of magnitude <= 127.
template <int S, bool HAD, bool VEC, int NCH, bool RAW>
__global__ void __launch_bounds__(kThreads, 2)
residual_mma_kernel(const double* __restrict__ X, int64_t rows, int64_t n, int64_t ldx,
                    const double* __restrict__ Qt, int rank, int64_t ldq, const double* __restrict__ coef,
                    const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_count,
                    int nblk, int ksplit, uint64_t cell_offset,
                    const double* __restrict__ inv_quantum, double* __restrict__ part_sumsq,
                    double* __restrict__ part_amax, int8_t* __restrict__ slices, int64_t rows_alloc,
                    int64_t n_pad, unsigned long long* __restrict__ energy_max, int prefetch,
                    int* __restrict__ exact_status) {
    // covariate block of the current 128 cells, shared by the CTA's 8 warps (they walk the same
    // cells): [buffer][covariate][cell], pitch 132 so that a B-fragment read (4 covariates x 8 cells
    // per half-warp) touches 16 distinct banks.  With 16-byte aligned operands (VEC) the kernel is
    // software-pipelined with cp.async: while a warp works on block b, its 8 x 128 tile of X for block
    // b + 1 (one private 8.5 KB tile per warp, row pitch 136 doubles: conflict-free 128-bit reads) and the
    // CTA's Qt block for b + 1 are already on their way into shared memory.  Without that the 8 warps of
    // a CTA, kept in lockstep by the staging barrier, all wait for DRAM at the same time and then all
    // compute at the same time (ncu r01n: 43 % of DRAM peak, issue slots 32 % busy, and neither fewer
    // instructions nor an L2 prefetch changed the time).
    extern __shared__ __align__(16) double s_dyn[];
    double (*s_q)[16][kQPitch] = reinterpret_cast<double (*)[16][kQPitch]>(s_dyn);
    constexpr bool ASYNC = VEC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* tile = s_dyn + 2 * 16 * kQPitch + warp * (kRowsW * kXPitch);
    const int g = lane >> 2, t = lane & 3;
    const int64_t n_logical = row_list ? (int64_t)*row_count : rows;
    if ((int64_t)blockIdx.x * kWarps * kRowsW >= n_logical) return;        // whole CTA idle
    const int64_t l0 = ((int64_t)blockIdx.x * kWarps + warp) * kRowsW;
    const int64_t lrow = l0 + g;
    const int64_t my_row = lrow < n_logical ? (row_list ? (int64_t)row_list[lrow] : lrow) : -1;
    const double* xr = X + (my_row >= 0 ? my_row : 0) * ldx + 2 * t;
    const double* cf = coef + (my_row >= 0 ? my_row : 0) * rank;
    const int b_begin = (int)((int64_t)nblk * blockIdx.y / ksplit);
    const int b_end = (int)((int64_t)nblk * (blockIdx.y + 1) / ksplit);
    const int nblk_full = (int)(n / 128);
    const int ngroup = (rank + 15) / 16;                 // covariates in groups of 16 = 4 MMA k-chunks
    const double my_invq = RAW ? 1.0 : ((my_row >= 0 && inv_quantum) ? inv_quantum[my_row] * (HAD ? kHadScale : 1.0) : 0.0);
    bool inexact = false;

    // A fragments (-coef[row g][4 ch + t]) of the first covariate group stay in registers
    double ca[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const int c = 4 * ch + t;
        ca[ch] = (my_row >= 0 && c < rank) ? -cf[c] : 0.0;
    }
    double sumsq = 0.0;
    uint32_t energy[S];                                  // sum of squared digits per plane (this lane's cells)
#pragma unroll
    for (int sidx = 0; sidx < S; ++sidx) energy[sidx] = 0;
    int amax_hi = 0;                                     // max over the high words of |z'| (monotone)
    const double sgn1 = (t & 1) ? -1.0 : 1.0, sgn2 = (t & 2) ? -1.0 : 1.0;

    // asynchronous staging of block `b`: this warp's X tile (16 x 16-byte chunks per lane, 512 contiguous
    // bytes per warp instruction) and the CTA's share of the Qt block
    const double* xrow = X + (my_row >= 0 ? my_row : 0) * ldx;           // row g of this warp's tile
    auto stage_async = [&](int b) {
        const int64_t kk0 = (int64_t)b * 128;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int r = j >> 1, c16 = ((j & 1) << 5) + lane;            // tile row, 16-byte chunk within the row
            const double* base = (const double*)__shfl_sync(0xffffffffu, (unsigned long long)xrow, 4 * r);
            const bool ok = __shfl_sync(0xffffffffu, my_row >= 0 ? 1 : 0, 4 * r) != 0;
            if (ok) cp_async16(tile + r * kXPitch + 2 * c16, base + kk0 + 2 * c16);
        }
        const int nchunk = min(rank, 16) * 64;
        for (int i = threadIdx.x; i < nchunk; i += kThreads) {
            const int c = i >> 6, c16 = i & 63;
            cp_async16(&s_q[b & 1][c][2 * c16], Qt + (int64_t)c * ldq + kk0 + 2 * c16);
        }
        cp_async_commit();
    };
    if (ASYNC) {
        // rows of s_q beyond the rank are read as zeros and never written again
        for (int i = threadIdx.x; i < 2 * 16 * kQPitch; i += kThreads) s_dyn[i] = 0.0;
        __syncthreads();
        if (b_begin < b_end && b_begin < nblk_full) stage_async(b_begin);
    }

    for (int blk = b_begin; blk < b_end; ++blk) {
        const int64_t k0 = (int64_t)blk * 128;
        const bool full = blk < nblk_full;
        double v[32];
        const int buf = blk & 1;
        if (ASYNC && full) {
            cp_async_wait_all();
            __syncthreads();              // every warp's Qt chunks of this block have landed (and block - 1 is done with)
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const double2 c2 = *reinterpret_cast<const double2*>(tile + g * kXPitch + 8 * u + 2 * t);
                v[2 * u] = c2.x; v[2 * u + 1] = c2.y;
            }
            __syncwarp();                 // the tile is in registers: refill it for the next block
            if (blk + 1 < b_end && blk + 1 < nblk_full) stage_async(blk + 1);
        } else {
            // ---- all X loads of the block first (independent, 8 KB per warp in flight)
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int64_t k = k0 + 8 * u + 2 * t;
                if (full) {
                    if (VEC) {
                        const double2 c2 = __ldg(reinterpret_cast<const double2*>(xr + k0 + 8 * u));
                        v[2 * u] = c2.x; v[2 * u + 1] = c2.y;
                    } else {
                        v[2 * u] = __ldg(xr + k0 + 8 * u);
                        v[2 * u + 1] = __ldg(xr + k0 + 8 * u + 1);
                    }
                } else {
                    v[2 * u] = (k < n) ? __ldg(xr + k0 + 8 * u) : 0.0;
                    v[2 * u + 1] = (k + 1 < n) ? __ldg(xr + k0 + 8 * u + 1) : 0.0;
                }
            }
            // ---- stage the first 16 covariates of this block in shared memory (coalesced rows)
            if (ASYNC) __syncthreads();   // (ragged last block after pipelined ones: the buffer may still be read)
            for (int i = threadIdx.x; i < 4 * NCH * 128; i += kThreads) {
                const int c = i >> 7, cell = i & 127;
                s_q[buf][c][cell] = (c < rank && (full || k0 + cell < n)) ? __ldg(Qt + (int64_t)c * ldq + k0 + cell) : 0.0;
            }
            __syncthreads();          // one barrier per block is enough with two buffers
        }
        // ---- residual tiles: C = X (8 rows x 8 cells), A = -coef, B = Qt
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            double c0v = v[2 * u], c1v = v[2 * u + 1];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch)               // rows >= rank of s_q hold zeros
                dmma(c0v, c1v, ca[ch], s_q[buf][4 * ch + t][8 * u + g]);
            if (NCH == 4 && ngroup > 1) {                              // rank > 16: further groups straight from L1/L2
                const int64_t kq = k0 + 8 * u + g;
                const bool qin = full || kq < n;
                for (int gq = 1; gq < ngroup; ++gq) {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const int cbase = 16 * gq + 4 * ch;
                        if (cbase < rank) {
                            const int c = cbase + t;
                            const double a = (my_row >= 0 && c < rank) ? -cf[c] : 0.0;
                            const double b = (qin && c < rank) ? __ldg(Qt + (int64_t)c * ldq + kq) : 0.0;
                            dmma(c0v, c1v, a, b);
                        }
                    }
                }
            }
            sumsq = fma(c0v, c0v, sumsq);
            sumsq = fma(c1v, c1v, sumsq);
            if (RAW) { c0v = v[2 * u]; c1v = v[2 * u + 1]; }      // the plane carries the raw row
            v[2 * u] = c0v;
            v[2 * u + 1] = c1v;
        }
        // ---- Walsh-Hadamard over the 128 cells: local index bits 0..4 in registers, t by shuffle
        if (HAD) {
            // first stage with the sign pattern folded in (compile-time negations of the operands)
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const double a = reg_flip(i) ? -v[i] : v[i], b = reg_flip(i + 1) ? -v[i + 1] : v[i + 1];
                v[i] = a + b;
                v[i + 1] = a - b;
            }
#pragma unroll
            for (int h = 2; h < 32; h <<= 1)
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if ((i & h) == 0) {
                        const double a = v[i], b = v[i + h];
                        v[i] = a + b;
                        v[i + h] = a - b;
                    }
#pragma unroll
            for (int i = 0; i < 32; ++i)                  // lower lane: v + p, upper lane: p - v
                v[i] = fma(sgn1, v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
#pragma unroll
            for (int i = 0; i < 32; ++i)
                v[i] = fma(sgn2, v[i], __shfl_xor_sync(0xffffffffu, v[i], 2));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) amax_hi = max(amax_hi, __double2hiint(v[i]) & 0x7fffffff);
        if (RAW) {
#pragma unroll
            for (int i = 0; i < 32; ++i) inexact |= !(fabs(v[i]) <= 127.0) || v[i] != rint(v[i]);
        }
        if (slices != nullptr && my_row >= 0)
            store_digits<S>(v, my_invq, slices + my_row * n_pad + k0 + 32 * t, rows_alloc * n_pad, energy);
    }
    if (RAW && inexact && my_row >= 0 && exact_status != nullptr) atomicOr(exact_status, 1);
    if (energy_max != nullptr && slices != nullptr) {
        // per-plane digit energy of each row over this CTA's cells -> maximum over rows, kept per
        // cell split: the host bounds every int32 partial sum of the contraction with Cauchy-Schwarz
#pragma unroll
        for (int sidx = 0; sidx < S; ++sidx) {
            double e = my_row >= 0 ? (double)energy[sidx] : 0.0;
            e += __shfl_xor_sync(0xffffffffu, e, 1);
            e += __shfl_xor_sync(0xffffffffu, e, 2);                         // row total (quad)
#pragma unroll
            for (int m = 4; m < 32; m <<= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, m));   // max over the 8 rows
            if (lane == 0) atomicMax(energy_max + blockIdx.y * NSR_MAX_SLICES + sidx, (unsigned long long)__double_as_longlong(e));
        }
    }
    if (part_sumsq != nullptr) {
        sumsq += __shfl_xor_sync(0xffffffffu, sumsq, 1);
        sumsq += __shfl_xor_sync(0xffffffffu, sumsq, 2);
        // upper bound of max|z'| from its high word (relative slack 2^-20): conservative for the
        // overflow test and costs the re-quantised rows one millionth of their scale
        double amax = amax_hi ? __hiloint2double(amax_hi + 1, 0) : 0.0;
        amax *= (HAD ? kHadScale : 1.0);
        amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        if (t == 0 && my_row >= 0) {
            part_sumsq[(int64_t)blockIdx.y * rows + my_row] = sumsq;
            part_amax[(int64_t)blockIdx.y * rows + my_row] = amax;
        }
    }
}

// var, final quantum, and the list of rows whose digits overflowed the estimated scale
__global__ void stats_finalize_kernel(const double* __restrict__ part_sumsq,
                                      const double* __restrict__ part_amax, int64_t rows,
                                      int ksplit, int64_t n, double vmax,
                                      double* __restrict__ var, double* __restrict__ quantum,
                                      double* __restrict__ inv_quantum, int32_t* __restrict__ fix_list,
                                      int32_t* __restrict__ fix_count, unsigned long long* __restrict__ energy_max,
                                      double raw_quantum) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double s = 0.0, m = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) {
        s += part_sumsq[(int64_t)ks * rows + i];
        m = fmax(m, part_amax[(int64_t)ks * rows + i]);
    }
    // A NaN / Inf anywhere in the row (or in the covariate basis) makes the sum of squares non-finite.
    // The reference asserts finite results (association.py:252-255); here the row poisons entry (0, 0) of
    // the energy report with a NaN - as an integer its bit pattern beats every finite energy under
    // atomicMax - and the host, which reads the report before every contraction, raises.
    if (!isfinite(s) && energy_max != nullptr) atomicMax(energy_max, 0x7ff8000000000000ull);
    double v = s / (double)n;
    if (v == 0.0) v = 1.0;                       // association.py:231,233
    var[i] = v;
    if (raw_quantum > 0.0) { quantum[i] = raw_quantum; return; }      // exact single-plane rows: fixed scale, no fix-up
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

// position p of a 128-cell block holds transformed index 8u + 2t + e with p = 32t + 2u + e
__global__ void unslice_kernel(const int8_t* __restrict__ slices, int64_t rows, int64_t rows_alloc,
                               int64_t n_pad, int n_slices, const double* __restrict__ quantum,
                               double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * n_pad) return;
    const int64_t r = i / n_pad, k = i % n_pad;
    const int p = (int)(k & 127);
    const int m = 8 * ((p & 31) >> 1) + 2 * (p >> 5) + (p & 1);
    int64_t v = 0;
    for (int s = 0; s < n_slices; ++s) v = v * 256 + slices[((int64_t)s * rows_alloc + r) * n_pad + k];
    out[r * n_pad + (k - p) + m] = (double)v * quantum[r];     // natural (transformed-index) order
}


}  // namespace

int nsr_use_hadamard = 1;   // test hook (nsr_set_option)
int nsr_prefetch = 0;       // test hook: L2 prefetch of the next block in pass B (measured neutral: off)

extern "C" int64_t nsr_padded_cells(int64_t n) { return (n + NSR_KBLOCK - 1) / NSR_KBLOCK * NSR_KBLOCK; }

// number of contiguous cell ranges the projection kernels split a row into: a function of n only
extern "C" int nsr_cell_splits(int64_t n) {
    const int64_t nblk = (n + NSR_KBLOCK - 1) / NSR_KBLOCK;
    int64_t ks = (nblk + 15) / 16;
    if (ks > NSR_MAX_SPLITS) ks = NSR_MAX_SPLITS;
    if (ks < 1) ks = 1;
    return (int)ks;
}

// n_slices = 1 with exact_status != nullptr: the exact single-plane mode of nsr_residualize_exact
static int residualize_impl(nsr_ctx* ctx, uintptr_t stream, const double* X, int64_t rows,
                            int64_t n, int64_t ldx, const double* Qt, int rank, int64_t ldq,
                            int n_slices, int8_t* slices, int64_t rows_alloc, int64_t n_pad,
                            double* quantum, double* var, double* coef, double* energy_max, int* exact_status) {
    NSR_REQUIRE(ctx != nullptr, "nsr_residualize: null context");
    NSR_REQUIRE(rows > 0 && rows < (1ll << 31) && n > 0 && ldx >= n,
                "nsr_residualize: bad shape rows=%lld n=%lld ldx=%lld", (long long)rows, (long long)n,
                (long long)ldx);
    NSR_REQUIRE(rank >= 0 && rank <= NSR_MAX_RANK, "nsr_residualize: rank %d outside [0,%d]", rank,
                NSR_MAX_RANK);
    NSR_REQUIRE(rank == 0 || (Qt != nullptr && ldq >= n), "nsr_residualize: bad covariate basis");
    const bool exact = exact_status != nullptr;
    NSR_REQUIRE(exact ? n_slices == 1 : (n_slices == 3 || n_slices == 4), "nsr_residualize: n_slices %d (3 or 4)", n_slices);
    NSR_REQUIRE(n_pad == nsr_padded_cells(n) && rows_alloc >= rows,
                "nsr_residualize: n_pad/rows_alloc inconsistent");
    NSR_REQUIRE(((uintptr_t)slices & 15) == 0, "nsr_residualize: slices must be 16-byte aligned");
    NSR_REQUIRE(((uintptr_t)X & 7) == 0 && ((uintptr_t)Qt & 7) == 0, "nsr_residualize: inputs must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    NSR_CHECK(cudaSetDevice(ctx->device));

    // 16-byte vector loads need even leading dimensions and 16-byte aligned bases
    const int vec = ((uintptr_t)X % 16 == 0) && (ldx % 2 == 0) &&
                    (rank == 0 || (((uintptr_t)Qt % 16 == 0) && (ldq % 2 == 0)));
    const int nblk = (int)(n_pad / 128);
    // The split of the cell axis depends on n only (2048 cells per CTA, at most 64 splits), never on
    // the number of rows in the call: partial sums are then combined in the same order whether a
    // matrix is residualised whole, in row chunks, or sharded over GPUs - results stay bit-identical.
    const int ksplit = nsr_cell_splits(n);
    const int64_t groups_w = (rows + kWarps * kRowsW - 1) / (kWarps * kRowsW);   // CTAs of 8 warps x 8 rows
    const int64_t groups_s = (rows + kWarps - 1) / kWarps;
    const int ks_a = ksplit, ks_b = ksplit;
    const int rk = rank > 0 ? rank : 1;

    const bool had = nsr_use_hadamard != 0;
    const int nch = rank >= 16 ? 4 : (rank + 3) / 4;

    const int ks_stat = ks_b;

    // scratch (doubles): coef partials | sum-x^2 partials | sumsq partials | amax partials | inv_quantum
    //                    | coef (if the caller passed none) ; then int32: fix_count(+pad) | fix_list
    const size_t n_part = (size_t)ks_a * rows * rk;
    const size_t n_psq = (size_t)ks_a * rows;
    const size_t n_stat = (size_t)ks_stat * rows, n_coef = (size_t)rows * rk;
    const size_t n_dbl = n_part + n_psq + 2 * n_stat + rows + n_coef;
    const size_t n_i32 = (size_t)(rows + 4);
    void* scratch = nullptr;
    if (nsr_scratch(ctx, n_dbl * sizeof(double) + n_i32 * sizeof(int32_t), &scratch)) return 1;
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
        const dim3 grid((unsigned)groups_w, (unsigned)ks_a);
        for (int c0 = 0; c0 < rank; c0 += 16) {          // 16 covariates per launch
#define NSR_LAUNCH_A(NQ, V) coef_mma_kernel<NQ, V><<<grid, kThreads, 0, st>>>(X, rows, n, ldx, Qt, rank, ldq, c0, ks_a, partial, psq)
            if (rank - c0 <= 8) { if (vec) NSR_LAUNCH_A(1, true); else NSR_LAUNCH_A(1, false); }
            else { if (vec) NSR_LAUNCH_A(2, true); else NSR_LAUNCH_A(2, false); }
#undef NSR_LAUNCH_A
        }
    } else {
        sumsq_kernel<<<dim3((unsigned)groups_s, (unsigned)ks_a), kThreads, 0, st>>>(X, rows, n, ldx, ks_a, psq);
    }
    coef_finalize_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(partial, psq, rows, rank, ks_a, n, vmax,
                                                                       coef_buf, invq, exact_status);
    const dim3 gridb((unsigned)groups_w, (unsigned)ks_b);
#define NSR_LAUNCH_B(S_, H_, V_, N_, LIST, COUNT, PS, PA)                                                    \
    NSR_CHECK(cudaFuncSetAttribute((const void*)residual_mma_kernel<S_, H_, V_, N_, (S_ == 1)>,              \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, kResidualSmem));            \
    residual_mma_kernel<S_, H_, V_, N_, (S_ == 1)><<<gridb, kThreads, kResidualSmem, st>>>(                  \
        X, rows, n, ldx, Qt, rank, ldq, coef_buf, LIST, COUNT, nblk, ks_b, (uint64_t)0, invq, PS, PA, slices, \
        rows_alloc, n_pad, (unsigned long long*)energy_max, nsr_prefetch, exact_status)
#define NSR_LAUNCH_B_N(S_, H_, V_, LIST, COUNT, PS, PA)                                                      \
    do {                                                                                                     \
        switch (nch) {                                                                                       \
            case 0: NSR_LAUNCH_B(S_, H_, V_, 0, LIST, COUNT, PS, PA); break;                                  \
            case 1: NSR_LAUNCH_B(S_, H_, V_, 1, LIST, COUNT, PS, PA); break;                                  \
            case 2: NSR_LAUNCH_B(S_, H_, V_, 2, LIST, COUNT, PS, PA); break;                                  \
            case 3: NSR_LAUNCH_B(S_, H_, V_, 3, LIST, COUNT, PS, PA); break;                                  \
            default: NSR_LAUNCH_B(S_, H_, V_, 4, LIST, COUNT, PS, PA); break;                                 \
        }                                                                                                    \
    } while (0)
#define NSR_LAUNCH_B_ALL(LIST, COUNT, PS, PA)                                                                \
    do {                                                                                                     \
        if (n_slices == 1) {                                                                                 \
            if (had) { if (vec) NSR_LAUNCH_B_N(1, true, true, LIST, COUNT, PS, PA); else NSR_LAUNCH_B_N(1, true, false, LIST, COUNT, PS, PA); }     \
            else { if (vec) NSR_LAUNCH_B_N(1, false, true, LIST, COUNT, PS, PA); else NSR_LAUNCH_B_N(1, false, false, LIST, COUNT, PS, PA); }        \
        } else if (n_slices == 3) {                                                                          \
            if (had) { if (vec) NSR_LAUNCH_B_N(3, true, true, LIST, COUNT, PS, PA); else NSR_LAUNCH_B_N(3, true, false, LIST, COUNT, PS, PA); }     \
            else { if (vec) NSR_LAUNCH_B_N(3, false, true, LIST, COUNT, PS, PA); else NSR_LAUNCH_B_N(3, false, false, LIST, COUNT, PS, PA); }        \
        } else {                                                                                             \
            if (had) { if (vec) NSR_LAUNCH_B_N(4, true, true, LIST, COUNT, PS, PA); else NSR_LAUNCH_B_N(4, true, false, LIST, COUNT, PS, PA); }     \
            else { if (vec) NSR_LAUNCH_B_N(4, false, true, LIST, COUNT, PS, PA); else NSR_LAUNCH_B_N(4, false, false, LIST, COUNT, PS, PA); }        \
        }                                                                                                    \
    } while (0)
    NSR_LAUNCH_B_ALL(nullptr, nullptr, p_sumsq, p_amax);
    stats_finalize_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(p_sumsq, p_amax, rows, ks_stat, n, vmax, var,
                                                                        quantum, invq, fix_list, fix_count,
                                                                        (unsigned long long*)energy_max,
                                                                        exact ? (had ? kHadScale : 1.0) : 0.0);
    // sparse fix-up: warps beyond the (device-side) count exit at once
    if (!exact) NSR_LAUNCH_B_ALL(fix_list, fix_count, nullptr, nullptr);
#undef NSR_LAUNCH_B_ALL
#undef NSR_LAUNCH_B_N
#undef NSR_LAUNCH_B
    NSR_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int nsr_residualize(nsr_ctx* ctx, uintptr_t stream, const double* X, int64_t rows,
                               int64_t n, int64_t ldx, const double* Qt, int rank, int64_t ldq,
                               int n_slices, int8_t* slices, int64_t rows_alloc, int64_t n_pad,
                               double* quantum, double* var, double* coef, double* energy_max) {
    return residualize_impl(ctx, stream, X, rows, n, ldx, Qt, rank, ldq, n_slices, slices, rows_alloc, n_pad, quantum, var,
                            coef, energy_max, nullptr);
}

extern "C" int nsr_residualize_exact(nsr_ctx* ctx, uintptr_t stream, const double* X, int64_t rows,
                                     int64_t n, int64_t ldx, const double* Qt, int rank, int64_t ldq,
                                     int8_t* plane, int64_t rows_alloc, int64_t n_pad, double* quantum,
                                     double* var, double* coef, double* energy_max, int* status) {
    NSR_REQUIRE(status != nullptr, "nsr_residualize_exact: null status");
    return residualize_impl(ctx, stream, X, rows, n, ldx, Qt, rank, ldq, 1, plane, rows_alloc, n_pad, quantum, var, coef,
                            energy_max, status);
}

namespace {
// coef[row][c] = sum over cell splits (fixed order), sumsq[row] likewise
__global__ void coef_sum_kernel(const double* __restrict__ partial, const double* __restrict__ psq, int64_t rows,
                                int rank, int ksplit, double* __restrict__ coef, double* __restrict__ sumsq) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    for (int c = 0; c < rank; ++c) {
        double s = 0.0;
        for (int ks = 0; ks < ksplit; ++ks) s += partial[((int64_t)ks * rows + row) * rank + c];
        coef[row * rank + c] = s;
    }
    if (sumsq) {
        double sq = 0.0;
        for (int ks = 0; ks < ksplit; ++ks) sq += psq[(int64_t)ks * rows + row];
        sumsq[row] = sq;
    }
}
}  // namespace

// Pass A of the projection on its own: coef = X Q^T (rows x rank, row-major) and sumsq = sum_k x^2
// for ANY (rank x n) matrix Q (no orthonormality assumed).  One streaming read of X on the FP64
// tensor cores; partial sums over the n-only cell splits are combined in a fixed order.
extern "C" int nsr_project_coef(nsr_ctx* ctx, uintptr_t stream, const double* X, int64_t rows, int64_t n,
                                int64_t ldx, const double* Q, int rank, int64_t ldq, double* coef, double* sumsq) {
    NSR_REQUIRE(ctx != nullptr && X != nullptr, "nsr_project_coef: null argument");
    NSR_REQUIRE(rows > 0 && rows < (1ll << 31) && n > 0 && ldx >= n, "nsr_project_coef: bad shape rows=%lld n=%lld",
                (long long)rows, (long long)n);
    NSR_REQUIRE(rank >= 0 && rank <= NSR_MAX_RANK && (rank == 0 || (Q != nullptr && ldq >= n && coef != nullptr)),
                "nsr_project_coef: bad Q (rank %d)", rank);
    NSR_REQUIRE(rank > 0 || sumsq != nullptr, "nsr_project_coef: nothing to compute");
    cudaStream_t st = (cudaStream_t)stream;
    NSR_CHECK(cudaSetDevice(ctx->device));
    const int vec = ((uintptr_t)X % 16 == 0) && (ldx % 2 == 0) && (rank == 0 || (((uintptr_t)Q % 16 == 0) && (ldq % 2 == 0)));
    const int ksplit = nsr_cell_splits(n);
    const int64_t groups_w = (rows + kWarps * kRowsW - 1) / (kWarps * kRowsW);
    const int64_t groups_s = (rows + kWarps - 1) / kWarps;
    const int rk = rank > 0 ? rank : 1;
    const size_t n_part = (size_t)ksplit * rows * rk, n_psq = (size_t)ksplit * rows;
    void* scratch = nullptr;
    if (nsr_scratch(ctx, (n_part + n_psq) * sizeof(double), &scratch)) return 1;
    double* partial = (double*)scratch;
    double* psq = partial + n_part;
    if (rank > 0) {
        const dim3 grid((unsigned)groups_w, (unsigned)ksplit);
        for (int c0 = 0; c0 < rank; c0 += 16) {
#define NSR_LAUNCH_A(NQ, V) coef_mma_kernel<NQ, V><<<grid, kThreads, 0, st>>>(X, rows, n, ldx, Q, rank, ldq, c0, ksplit, partial, psq)
            if (rank - c0 <= 8) { if (vec) NSR_LAUNCH_A(1, true); else NSR_LAUNCH_A(1, false); }
            else { if (vec) NSR_LAUNCH_A(2, true); else NSR_LAUNCH_A(2, false); }
#undef NSR_LAUNCH_A
        }
    } else {
        sumsq_kernel<<<dim3((unsigned)groups_s, (unsigned)ksplit), kThreads, 0, st>>>(X, rows, n, ldx, ksplit, psq);
    }
    coef_sum_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(partial, psq, rows, rank, ksplit, coef, sumsq);
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
