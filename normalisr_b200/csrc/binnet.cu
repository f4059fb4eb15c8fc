// P-value network -> binary network: per-row Benjamini-Hochberg + threshold, without sorting.
// Reference: src/normalisr/binnet.py:134-170 (binnet) calling bh (:77-131) on every row with the
// diagonal removed.  HBM-streaming: 8 B read + 1 B written per matrix entry.
//
// bh() gives entry j of a row (n0 off-diagonal entries, c_j = #{k : p_k <= p_j}) the Q-value
//   q_j = min_{k : p_k >= p_j} min(1, p_k / (c_k / n0)),
// so for qcut < 1:  q_j <= qcut  <=>  p_j <= p_(k*),  k* = max{c_k : p_k / (c_k / n0) <= qcut}.
// With g(c) = #{k : p_k / (c / n0) <= qcut} (non-increasing as c decreases) the sequence
// c <- g(c) started at c = n0 decreases to the largest fixed point, which is k*: every index
// that passes the reference's test at its own rank also passes it at any larger c, so the
// iterate never drops below k*, and a fixed point c = g(c) is itself an index set that passes.
// The kernel evaluates the SAME floating-point expression as the reference, p / (c / n0) <= qcut
// (binnet.py:122-124 computes w = cumsum / n0 and p / w), so the boolean output is bit-identical.
// One CTA per row: the row is read from HBM once into shared memory (rows up to 28,000 entries;
// wider rows re-read themselves from L2) and the iteration and the final threshold run from there.
#include "nsr_common.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr int kSmemRowMax = 28000;     // doubles of a row kept in shared memory (224 KB of the 227 KB)

__device__ __forceinline__ int block_sum(int v, int* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                   // s_red reuse
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
    return t;
}

// Largest double t with fl(t / w) <= qcut.  Correctly rounded division is monotone in its
// numerator, so  fl(p / w) <= qcut  <=>  p <= t : the row passes compare against t instead of
// dividing every entry.  fl(qcut * w) is within a few ulp of t; step to it exactly.
__device__ __forceinline__ double bh_threshold(double w, double qcut) {
    double t = qcut * w;
    while (t / w > qcut) t = nextafter(t, 0.0);
    for (;;) {
        const double u = nextafter(t, 2.0);
        if (u / w <= qcut) t = u; else break;
    }
    return t;
}

__device__ __forceinline__ uint32_t bn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// entries <= thr among s_row[0, cols): two per 16-byte shared-memory load
__device__ __forceinline__ int count_le(const double* s_row, int64_t cols, double thr) {
    int mine = 0;
    const double2* v = reinterpret_cast<const double2*>(s_row);
    const int64_t half = cols >> 1;
#pragma unroll 4
    for (int64_t j = threadIdx.x; j < half; j += kThreads) {
        const double2 p = v[j];
        mine += (p.x <= thr ? 1 : 0) + (p.y <= thr ? 1 : 0);
    }
    if ((cols & 1) && threadIdx.x == 0) mine += s_row[cols - 1] <= thr ? 1 : 0;
    return mine;
}

// SMEM = true: the row is staged once in shared memory (one bulk asynchronous copy when the row is
// 16-byte aligned, plain loads otherwise) and every later pass reads it from there; false: rows
// wider than shared memory re-read themselves from L2.
template <bool SMEM>
__global__ void __launch_bounds__(kThreads)
binnet_rows_kernel(const double* __restrict__ P, int64_t cols, int64_t ld, int64_t diag0, double qcut,
                   uint8_t* __restrict__ net, int64_t ld_net, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) double s_row[];
    __shared__ int s_red[kThreads / 32];
    __shared__ __align__(8) uint64_t s_bar;
    const int64_t row = blockIdx.x;
    const double* p_row = P + row * ld;
    const int64_t diag = row + diag0;                  // column of this row's diagonal entry
    const bool has_diag = diag >= 0 && diag < cols;
    const int64_t n0 = cols - (has_diag ? 1 : 0);
    int bad = 0, mine = 0;

    if (SMEM) {
        const bool bulk = (((uintptr_t)p_row & 15) == 0) && ((cols & 1) == 0);
        if (bulk) {
            // one thread posts the whole row (cols * 8 bytes) as a single bulk copy; all wait on the mbarrier
            const uint32_t bar = bn_smem_u32(&s_bar);
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const uint32_t bytes = (uint32_t)(cols * sizeof(double));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(bn_smem_u32(s_row)), "l"(p_row), "r"(bytes), "r"(bar) : "memory");
            }
            uint32_t ok = 0;
            const long long t0 = clock64();
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
                if (!ok && clock64() - t0 > 4000000000ll) __trap();      // never hang the device
            }
        } else {
#pragma unroll 8
            for (int64_t j = threadIdx.x; j < cols; j += kThreads) s_row[j] = p_row[j];
            __syncthreads();
        }
        // validate (binnet.py:152-153, the diagonal included), then make the diagonal entry inert
#pragma unroll 4
        for (int64_t j = threadIdx.x; j < cols; j += kThreads) {
            const double p = s_row[j];
            if (!(p >= 0.0 && p <= 1.0)) bad = 1;      // also catches NaN
        }
        __syncthreads();
        if (has_diag && threadIdx.x == 0) s_row[diag] = 2.0;
        __syncthreads();
        mine = count_le(s_row, cols, qcut);            // only p <= qcut can pass at all (c / n0 <= 1)
    } else {
#pragma unroll 8
        for (int64_t j = threadIdx.x; j < cols; j += kThreads) {
            const double p = p_row[j];
            if (!(p >= 0.0 && p <= 1.0)) bad = 1;
            mine += (j != diag && p <= qcut) ? 1 : 0;
        }
    }
    if (bad) atomicAdd(&stats[1], 1ull);
    int c = block_sum(mine, s_red);                    // g(n0): w = 1
    const double n0d = (double)n0;
    double thr = qcut;
    while (c > 0) {
        thr = bh_threshold((double)c / n0d, qcut);
        if (SMEM) {
            mine = count_le(s_row, cols, thr);
        } else {
            mine = 0;
#pragma unroll 8
            for (int64_t j = threadIdx.x; j < cols; j += kThreads) mine += (j != diag && p_row[j] <= thr) ? 1 : 0;
        }
        const int c_new = block_sum(mine, s_red);
        if (c_new == c) break;
        c = c_new;
    }
    // pass 2: threshold with the final rank
    uint8_t* o_row = net + row * ld_net;
    if (c == 0) thr = -1.0;
    if (SMEM) {
        int64_t done = 0;
        if (((uintptr_t)o_row & 7) == 0) {             // 8 entries -> one 8-byte store
            const double2* v = reinterpret_cast<const double2*>(s_row);
            const int64_t oct = cols >> 3;
            for (int64_t j = threadIdx.x; j < oct; j += kThreads) {
                uint64_t bits = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double2 p = v[4 * j + q];
                    bits |= (uint64_t)(p.x <= thr ? 1 : 0) << (16 * q);
                    bits |= (uint64_t)(p.y <= thr ? 1 : 0) << (16 * q + 8);
                }
                reinterpret_cast<uint64_t*>(o_row)[j] = bits;
            }
            done = oct << 3;
        }
        for (int64_t j = done + threadIdx.x; j < cols; j += kThreads) o_row[j] = (s_row[j] <= thr) ? 1 : 0;
    } else {
#pragma unroll 8
        for (int64_t j = threadIdx.x; j < cols; j += kThreads) o_row[j] = (j != diag && p_row[j] <= thr) ? 1 : 0;
    }
    if (c > 0 && threadIdx.x == 0) atomicAdd(&stats[0], (unsigned long long)c);
}

}  // namespace

extern "C" int nsr_binnet(nsr_ctx* ctx, uintptr_t stream, const double* P, int64_t rows, int64_t cols, int64_t ld,
                          int64_t diag0, double qcut, uint8_t* net, int64_t ld_net, unsigned long long* stats) {
    NSR_REQUIRE(ctx && P && net && stats, "nsr_binnet: null argument");
    NSR_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols && ld_net >= cols && rows < (1ll << 31),
                "nsr_binnet: bad shape rows=%lld cols=%lld", (long long)rows, (long long)cols);
    NSR_REQUIRE(qcut > 0.0 && qcut < 1.0, "nsr_binnet: qcut must be in (0, 1)");
    NSR_CHECK(cudaSetDevice(ctx->device));
    if (cols <= kSmemRowMax) {
        const int smem = (int)(cols * sizeof(double));
        NSR_CHECK(cudaFuncSetAttribute(binnet_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kSmemRowMax * (int)sizeof(double)));
        binnet_rows_kernel<true><<<(unsigned)rows, kThreads, smem, (cudaStream_t)stream>>>(P, cols, ld, diag0, qcut, net,
                                                                                            ld_net, stats);
    } else {
        binnet_rows_kernel<false><<<(unsigned)rows, kThreads, 0, (cudaStream_t)stream>>>(P, cols, ld, diag0, qcut, net,
                                                                                          ld_net, stats);
    }
    NSR_CHECK(cudaGetLastError());
    return 0;
}
