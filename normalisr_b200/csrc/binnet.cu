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
// Only entries with p <= qcut can ever pass (c / n0 <= 1), so the iteration runs over those
// candidates, compacted into shared memory while the row streams in (a row with more candidates
// than fit re-reads itself from L2 instead).
#include "nsr_common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kCandCap = 6000;     // candidates kept in shared memory (47 KB, static limit 48 KB)

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

__global__ void __launch_bounds__(kThreads)
binnet_rows_kernel(const double* __restrict__ P, int64_t cols, int64_t ld, int64_t diag0, double qcut,
                   uint8_t* __restrict__ net, int64_t ld_net, unsigned long long* __restrict__ stats) {
    __shared__ double s_cand[kCandCap];
    __shared__ int s_red[kThreads / 32];
    __shared__ int s_count;
    const int64_t row = blockIdx.x;
    const double* p_row = P + row * ld;
    const int64_t diag = row + diag0;                  // column of this row's diagonal entry
    const int64_t n0 = cols - ((diag >= 0 && diag < cols) ? 1 : 0);
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();

    // pass 1: stream the row, validate, compact the candidates
    int bad = 0;
    for (int64_t j = threadIdx.x; j < cols; j += kThreads) {
        const double p = p_row[j];
        if (!(p >= 0.0 && p <= 1.0)) bad = 1;          // also catches NaN (binnet.py:152-153)
        if (j != diag && p <= qcut) {
            const int slot = atomicAdd(&s_count, 1);
            if (slot < kCandCap) s_cand[slot] = p;
        }
    }
    if (bad) atomicAdd(&stats[1], 1ull);
    __syncthreads();
    int c = s_count;                                   // g(n0): w = n0 / n0 = 1, p / 1 <= qcut
    const bool in_smem = c <= kCandCap;
    const double n0d = (double)n0;
    double w = 1.0;
    while (c > 0) {
        w = (double)c / n0d;
        int mine = 0;
        if (in_smem) {
            for (int k = threadIdx.x; k < s_count; k += kThreads) mine += (s_cand[k] / w <= qcut) ? 1 : 0;
        } else {
            for (int64_t j = threadIdx.x; j < cols; j += kThreads)
                mine += (j != diag && p_row[j] / w <= qcut) ? 1 : 0;
        }
        const int c_new = block_sum(mine, s_red);
        if (c_new == c) break;
        c = c_new;
    }
    // pass 2: the row again (L2), threshold with the final w
    uint8_t* o_row = net + row * ld_net;
    if (c == 0) {
        for (int64_t j = threadIdx.x; j < cols; j += kThreads) o_row[j] = 0;
    } else {
        for (int64_t j = threadIdx.x; j < cols; j += kThreads)
            o_row[j] = (j != diag && p_row[j] / w <= qcut) ? 1 : 0;
        if (threadIdx.x == 0) atomicAdd(&stats[0], (unsigned long long)c);
    }
}

}  // namespace

extern "C" int nsr_binnet(nsr_ctx* ctx, uintptr_t stream, const double* P, int64_t rows, int64_t cols, int64_t ld,
                          int64_t diag0, double qcut, uint8_t* net, int64_t ld_net, unsigned long long* stats) {
    NSR_REQUIRE(ctx && P && net && stats, "nsr_binnet: null argument");
    NSR_REQUIRE(rows >= 1 && cols >= 1 && ld >= cols && ld_net >= cols && rows < (1ll << 31),
                "nsr_binnet: bad shape rows=%lld cols=%lld", (long long)rows, (long long)cols);
    NSR_REQUIRE(qcut > 0.0 && qcut < 1.0, "nsr_binnet: qcut must be in (0, 1)");
    NSR_CHECK(cudaSetDevice(ctx->device));
    binnet_rows_kernel<<<(unsigned)rows, kThreads, 0, (cudaStream_t)stream>>>(P, cols, ld, diag0, qcut, net, ld_net, stats);
    NSR_CHECK(cudaGetLastError());
    return 0;
}
