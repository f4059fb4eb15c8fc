// Shared declarations for the normalisr_b200 CUDA library (internal).
#pragma once
#define NSR_TILE_COUNTERS 64
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/normalisr_b200.h"

struct nsr_ctx {
    int device = 0;
    int sm_count = 0;
    // grow-only scratch owned by the context (projection partials, tile lists)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    int32_t* tiles_dev = nullptr;
    int32_t* tiles_pinned = nullptr;   // pinned staging of the tile list (same capacity), guarded by tiles_event
    cudaEvent_t tiles_event = nullptr;
    size_t tiles_cap = 0;
    void* encode_tiled = nullptr;      // cuTensorMapEncodeTiled, resolved at create
    // dynamic tile scheduler of the tcgen05 kernel: one counter per launch, used round-robin so
    // launches in flight on different streams never share one
    int* tile_counters = nullptr;
    unsigned launch_seq = 0;
    // adaptive schedule: per-tile flags | compacted tile list | count
    int32_t* refine_dev = nullptr;
    size_t refine_cap = 0;
};

void nsr_set_error(const char* fmt, ...);
int nsr_scratch(nsr_ctx* ctx, size_t bytes, void** out);

#define NSR_CHECK(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            nsr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                          cudaGetErrorString(_e));                                   \
            return 1;                                                                \
        }                                                                            \
    } while (0)

#define NSR_REQUIRE(cond, ...)                                                       \
    do {                                                                             \
        if (!(cond)) {                                                               \
            nsr_set_error(__VA_ARGS__);                                              \
            return 2;                                                                \
        }                                                                            \
    } while (0)

// ---- digit slicing ---------------------------------------------------------------
// A residual value is stored as round(z / quantum) written in balanced base 256:
//   V = sum_{a=1..S} d_a 256^(S-a),  d_a in [-128, 127]  (d_1 in [-127, 127]).
// quantum = absmax / NSR_VMAX(S) keeps |V| <= 127 * 256^(S-1) so d_1 never overflows.
#if defined(__CUDACC__)
#define NSR_HDI __host__ __device__ __forceinline__
#else
#define NSR_HDI inline
#endif

NSR_HDI double nsr_vmax(int n_slices) { return 127.0 * (double)(1ll << (8 * (n_slices - 1))); }

// digits[0] is the most significant one
NSR_HDI void nsr_digits(int32_t v, int n_slices, int8_t* digits) {
    for (int a = n_slices - 1; a >= 1; --a) {
        int8_t d = (int8_t)(v & 0xFF);     // low byte, sign-extended: balanced digit
        digits[a] = d;
        v = (v - (int32_t)d) >> 8;         // exact
    }
    digits[0] = (int8_t)v;
}

// ---- products kept by the contraction -------------------------------------------------
// All digit pairs (a, b) with a + b <= wmax are multiplied; pairs with equal a + b share an
// int32 accumulator ("weight group" g = a + b - 2).
//   (S=3, products=6): wmax = 4, 3 groups     (S=3, products=8): wmax = 5, 4 groups
//   (S=4, products=10): wmax = 5, 4 groups
NSR_HDI int nsr_wmax(int n_slices, int n_products) {
    if (n_slices == 3 && n_products == 6) return 4;
    if (n_slices == 3 && n_products == 8) return 5;
    if (n_slices == 4 && n_products == 10) return 5;
    if (n_slices == 2 && n_products == 3) return 3;
    return -1;
}
// A operand with ONE exact plane (nsr_residualize_exact) against sb planes of B: every product is kept
//   (1, 1): 1 product, wmax 2     (1, 3): 3 products, wmax 4     (1, 4): 4 products, wmax 5
NSR_HDI int nsr_wmax_ab(int sa, int sb, int n_products) {
    if (sa == sb) return nsr_wmax(sa, n_products) > 0 ? nsr_wmax(sa, n_products) : ((sa == 1 && n_products == 1) ? 2 : -1);
    if (sa == 1 && n_products == sb && (sb == 3 || sb == 4)) return sb + 1;
    return -1;
}
