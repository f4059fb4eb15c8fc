// CUDA-core (dp4a) contraction: same integer arithmetic as the tcgen05 kernel, trivially
// auditable.  It exists so the tests can demand BIT-IDENTICAL int32 sums from the tensor-core
// kernel at sizes the CPU oracle cannot reach.  Not the product path (NSR_ENGINE_SIMT).
#include "epilogue.cuh"

namespace {

constexpr int kSub = 64;          // a CTA computes a 64x64 quarter of a 128x128 tile
constexpr int kKc = 64;           // cells staged per step
constexpr int kPitchW = kKc / 4 + 1;   // smem row pitch in 32-bit words (+1: bank spread)

struct SimtArgs {
    const int8_t* a; const int8_t* b;
    int64_t rows_alloc_a, rows_alloc_b, n_pad, cell_begin, cell_end;
    int sa, sb, wmax;
    const int32_t* tiles;
    ContractParams ep;
};

__global__ void __launch_bounds__(256) contract_simt_kernel(const SimtArgs g) {
    __shared__ uint32_t sa[NSR_MAX_SLICES][kSub][kPitchW];
    __shared__ uint32_t sb[NSR_MAX_SLICES][kSub][kPitchW];
    const int tile_r = g.tiles[2 * blockIdx.x], tile_c = g.tiles[2 * blockIdx.x + 1];
    const int64_t row0 = (int64_t)tile_r * NSR_TILE + (blockIdx.y >> 1) * kSub;
    const int64_t col0 = (int64_t)tile_c * NSR_TILE + (blockIdx.y & 1) * kSub;
    if (row0 >= g.ep.rows_a || col0 >= g.ep.rows_b) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int SA = g.sa, SB = g.sb;

    int32_t acc[4][4][4];
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[w][r][c] = 0;

    const int lrow = threadIdx.x >> 2, lchunk = threadIdx.x & 3;    // 64 rows x 4 x 16 B
    for (int64_t k0 = g.cell_begin; k0 < g.cell_end; k0 += kKc) {
        for (int s = 0; s < (SA > SB ? SA : SB); ++s) {
            uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
            if (s < SA && row0 + lrow < g.ep.rows_a)
                va = *reinterpret_cast<const uint4*>(g.a + ((int64_t)s * g.rows_alloc_a + row0 + lrow) * g.n_pad + k0 + 16 * lchunk);
            if (s < SB && col0 + lrow < g.ep.rows_b)
                vb = *reinterpret_cast<const uint4*>(g.b + ((int64_t)s * g.rows_alloc_b + col0 + lrow) * g.n_pad + k0 + 16 * lchunk);
            uint32_t* da = &sa[s][lrow][4 * lchunk];
            uint32_t* db = &sb[s][lrow][4 * lchunk];
            da[0] = va.x; da[1] = va.y; da[2] = va.z; da[3] = va.w;
            db[0] = vb.x; db[1] = vb.y; db[2] = vb.z; db[3] = vb.w;
        }
        __syncthreads();
#pragma unroll 2
        for (int kw = 0; kw < kKc / 4; ++kw) {
            int32_t av[NSR_MAX_SLICES][4], bv[NSR_MAX_SLICES][4];
#pragma unroll
            for (int s = 0; s < NSR_MAX_SLICES; ++s)
                {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        av[s][r] = s < SA ? (int32_t)sa[s][4 * ty + r][kw] : 0;
                        bv[s][r] = s < SB ? (int32_t)sb[s][4 * tx + r][kw] : 0;
                    }
                }
#pragma unroll
            for (int da = 0; da < NSR_MAX_SLICES; ++da)
#pragma unroll
                for (int db = 0; db < NSR_MAX_SLICES; ++db) {
                    const int w = da + db;               // (a-1)+(b-1) = weight group
                    if (da < SA && db < SB && w < 4 && w + 2 <= g.wmax) {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                acc[w][r][c] = __dp4a(av[da][r], bv[db][c], acc[w][r][c]);
                    }
                }
        }
        __syncthreads();
    }
    const bool mirror = g.ep.mode == NSR_MODE_COEX && tile_r != tile_c;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t i = row0 + 4 * ty + r;
        if (i >= g.ep.rows_a) continue;
        const double qi = g.ep.qa[i], vi = g.ep.va ? g.ep.va[i] : 1.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t j = col0 + 4 * tx + c;
            if (j >= g.ep.rows_b) continue;
            int32_t a4[4] = {acc[0][r][c], acc[1][r][c], acc[2][r][c], acc[3][r][c]};
            nsr_finish(g.ep, i, j, qi, vi, g.ep.qb[j], g.ep.vb ? g.ep.vb[j] : 1.0,
                       nsr_combine(g.ep, a4), mirror);
        }
    }
}

}  // namespace

int nsr_launch_contract_simt(cudaStream_t st, const int8_t* a, int64_t rows_alloc_a, int n_slices_a, const int8_t* b,
                             int64_t rows_alloc_b, int n_slices_b, int64_t n_pad, int wmax,
                             const int32_t* tiles_dev, int64_t n_tiles, const ContractParams& ep,
                             int64_t cell_begin, int64_t cell_end) {
    SimtArgs g;
    g.cell_begin = cell_begin; g.cell_end = cell_end;
    g.a = a; g.b = b; g.rows_alloc_a = rows_alloc_a; g.rows_alloc_b = rows_alloc_b; g.n_pad = n_pad;
    g.sa = n_slices_a; g.sb = n_slices_b; g.wmax = wmax; g.tiles = tiles_dev; g.ep = ep;
    contract_simt_kernel<<<dim3((unsigned)n_tiles, 4), 256, 0, st>>>(g);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
