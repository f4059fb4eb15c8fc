// C ABI of normalisr_b200 (see include/normalisr_b200.h for the contract of each entry).
#include <stdarg.h>
#include <string.h>

#include <unordered_map>
#include <vector>

#include <cuda.h>

#include "epilogue.cuh"

int nsr_launch_contract_simt(cudaStream_t st, const int8_t* a, int64_t rows_alloc_a, int n_slices_a, const int8_t* b,
                             int64_t rows_alloc_b, int n_slices_b, int64_t n_pad, int wmax,
                             const int32_t* tiles_dev, int64_t n_tiles, const ContractParams& ep,
                             int64_t cell_begin, int64_t cell_end);
int nsr_launch_contract_umma(nsr_ctx* ctx, cudaStream_t st, const int8_t* a, int64_t rows_a,
                             int64_t rows_alloc_a, int n_slices_a, const NsrSegOperand* segs, int n_segs,
                             int64_t n_pad, int n_slices_b, int wmax,
                             const int32_t* tiles_dev, int64_t n_tiles, const ContractParams& ep,
                             int64_t cell_begin, int64_t cell_end, const int* n_tiles_dev, int n_parts = 1,
                             double* part_out = nullptr, int64_t part_stride = 0);
int nsr_launch_contract_finish(cudaStream_t st, const int32_t* tiles_dev, int64_t n_tiles, const ContractParams& ep,
                               const SegInfo& sg, int n_parts, const double* part, int64_t part_stride);
int nsr_umma_parts(int64_t cells, int n_slices_a, int n_slices_b, int n_parts);
extern int nsr_use_hadamard;
extern int nsr_prefetch;
extern int nsr_umma_kblock;
extern int nsr_umma_pair;
extern int nsr_epi_warps;
extern int nsr_umma_stack;
extern int nsr_binnet_keys;
extern int nsr_umma_dynamic;
int nsr_split_k = 1;          // 1: few-tile launches split the cells over work items (see contract_impl)
extern int nsr_epi_sleep_ns;

// Adaptive digit-product schedule, OPT-IN (tcgen05 engine, 3 planes / 8 products, one pass over the
// cells): for n >= nsr_adaptive_min_cells every tile first runs the 6-product schedule (weight-5 products
// (2,3), (3,2) dropped: |dr| ~ 1.0e-6 / sqrt(n) at random sign, harmless unless the pair is extremely
// significant), tiles holding a pair with r^2 n > kRefineZ2 are redone with all 8.  For an
// unrefined pair the relative error of P is n |r| dr <= sqrt(kRefineZ2) * 1.0e-6 = 8e-6 (1 sigma).
// It pays only when extremely significant pairs are confined to few tiles (screens, sparse networks):
// on the 100k x 20k benchmark matrix 86 % of the tiles hold such a pair and the two phases cost twice
// the full schedule (profiles/r01am_adaptive_ab.md), hence off unless asked for.
int nsr_adaptive_min_cells = 0;        // 0 = always the full schedule; >= 8192 sensible when enabled
static constexpr double kRefineZ2 = 64.0;

namespace {
// tiles flagged by the first phase -> compact list (order irrelevant: integer sums) + count
__global__ void refine_compact_kernel(const int* __restrict__ need, const int32_t* __restrict__ tiles, int n_tiles,
                                      int32_t* __restrict__ out_tiles, int* __restrict__ out_count) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += gridDim.x * blockDim.x) {
        if (need[t]) {
            const int slot = atomicAdd(out_count, 1);
            out_tiles[2 * slot] = tiles[2 * t];
            out_tiles[2 * slot + 1] = tiles[2 * t + 1];
        }
    }
}
}  // namespace

static thread_local char g_err[1024] = "";

void nsr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int nsr_scratch(nsr_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_bytes) {
        // grow-only; a synchronising free is fine here (size changes are rare)
        if (ctx->scratch) NSR_CHECK(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes + bytes / 4;
        NSR_CHECK(cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return 0;
}

extern "C" int nsr_version(void) { return NSR_VERSION; }
extern "C" const char* nsr_last_error(void) { return g_err; }

extern "C" int nsr_ctx_create(int device, nsr_ctx** out) {
    NSR_REQUIRE(out != nullptr, "nsr_ctx_create: null output pointer");
    int count = 0;
    NSR_CHECK(cudaGetDeviceCount(&count));
    NSR_REQUIRE(device >= 0 && device < count, "nsr_ctx_create: device %d of %d", device, count);
    NSR_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    NSR_CHECK(cudaGetDeviceProperties(&prop, device));
    NSR_REQUIRE(prop.major == 10, "normalisr_b200 is built for sm_100a only; device %d is sm_%d%d (%s)",
                device, prop.major, prop.minor, prop.name);
    nsr_ctx* ctx = new nsr_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) fn = nullptr;
    ctx->encode_tiled = fn;
    if (cudaMalloc(&ctx->tile_counters, NSR_TILE_COUNTERS * sizeof(int)) != cudaSuccess) {
        delete ctx;
        nsr_set_error("nsr_ctx_create: cudaMalloc failed");
        return 1;
    }
    *out = ctx;
    return 0;
}

extern "C" int nsr_ctx_destroy(nsr_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->tiles_dev) cudaFree(ctx->tiles_dev);
    if (ctx->tiles_pinned) cudaFreeHost(ctx->tiles_pinned);
    if (ctx->tiles_event) cudaEventDestroy(ctx->tiles_event);
    if (ctx->tile_counters) cudaFree(ctx->tile_counters);
    if (ctx->refine_dev) cudaFree(ctx->refine_dev);
    delete ctx;
    return 0;
}

// test hooks: "hadamard" (0/1), "umma_kblock" (64/128)
extern "C" int nsr_set_option(const char* name, int value) {
    if (!strcmp(name, "hadamard")) { nsr_use_hadamard = value ? 1 : 0; return 0; }
    if (!strcmp(name, "prefetch")) { nsr_prefetch = value ? 1 : 0; return 0; }
    if (!strcmp(name, "umma_pair")) { nsr_umma_pair = value ? 1 : 0; return 0; }
    if (!strcmp(name, "umma_dynamic")) { nsr_umma_dynamic = value ? 1 : 0; return 0; }
    if (!strcmp(name, "split_k")) { nsr_split_k = value ? 1 : 0; return 0; }
    if (!strcmp(name, "adaptive_min_cells")) { nsr_adaptive_min_cells = value < 0 ? 0 : value; return 0; }
    if (!strcmp(name, "binnet_keys")) { nsr_binnet_keys = value ? 1 : 0; return 0; }
    if (!strcmp(name, "umma_stack")) { nsr_umma_stack = value ? 1 : 0; return 0; }
    if (!strcmp(name, "epi_warps")) {
        NSR_REQUIRE(value == 8 || value == 16, "epi_warps must be 8 or 16");
        nsr_epi_warps = value;
        return 0;
    }
    if (!strcmp(name, "epi_sleep_ns")) { nsr_epi_sleep_ns = value < 0 ? 0 : value; return 0; }
    if (!strcmp(name, "umma_kblock")) {
        NSR_REQUIRE(value == 64 || value == 128, "umma_kblock must be 64 or 128");
        nsr_umma_kblock = value;
        return 0;
    }
    nsr_set_error("nsr_set_option: unknown option %s", name);
    return 2;
}

// Shared implementation of nsr_contract / nsr_contract_ab / nsr_contract_segments.
// tile_stride = 2: host_tiles holds (tile_row, tile_col), one segment; 3: (segment, tile_row, tile_col).
// tile list: page-locked host memory (read over PCIe, zero-copy) -> device memory
__global__ void tiles_upload_kernel(int4* __restrict__ dst, const int4* __restrict__ src, int64_t n_vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

static int contract_impl(nsr_ctx* ctx, uintptr_t stream, int engine, int mode, const int8_t* a_slices, int64_t rows_a,
                         int64_t rows_alloc_a, int n_slices_a, const double* quantum_a, const double* var_a,
                         const NsrSegOperand* segs, int n_segs, int n_slices_b, int64_t n, int64_t n_pad, int n_products,
                         const int32_t* host_tiles, int tile_stride, int64_t n_tiles, double dof_a, double* P,
                         double* out2, int64_t ld, int64_t k_chunk) {
    NSR_REQUIRE(ctx != nullptr, "nsr_contract: null context");
    NSR_REQUIRE(engine == NSR_ENGINE_UMMA || engine == NSR_ENGINE_SIMT, "nsr_contract: unknown engine %d", engine);
    NSR_REQUIRE(n_segs >= 1 && n_segs <= NSR_MAX_SEGMENTS && segs != nullptr, "nsr_contract: %d segments (1..%d)", n_segs,
                NSR_MAX_SEGMENTS);
    NSR_REQUIRE(n_segs == 1 || engine == NSR_ENGINE_UMMA, "nsr_contract: segments need the tcgen05 engine");
    NSR_REQUIRE(rows_a > 0 && n > 0 && n_pad == nsr_padded_cells(n), "nsr_contract: bad shape rows_a=%lld n=%lld n_pad=%lld",
                (long long)rows_a, (long long)n, (long long)n_pad);
    const int wmax = nsr_wmax_ab(n_slices_a, n_slices_b, n_products);
    NSR_REQUIRE(wmax > 0, "nsr_contract: unsupported (n_slices_a=%d, n_slices_b=%d, n_products=%d)", n_slices_a, n_slices_b,
                n_products);
    NSR_REQUIRE(out2 != nullptr && quantum_a && a_slices, "nsr_contract: null buffer");
    NSR_REQUIRE(((uintptr_t)a_slices & 15) == 0, "nsr_contract: slice planes must be 16-byte aligned");
    bool need_p = false;
    for (int s = 0; s < n_segs; ++s) {
        const SegInfo& si = segs[s].info;
        NSR_REQUIRE(si.mode == NSR_MODE_COEX || si.mode == NSR_MODE_DE || si.mode == NSR_MODE_RAW ||
                        si.mode == NSR_MODE_COEX_UPPER || si.mode == NSR_MODE_COEX_RECT,
                    "nsr_contract: unknown mode %d", si.mode);
        const bool sym = si.mode == NSR_MODE_COEX || si.mode == NSR_MODE_COEX_UPPER;
        NSR_REQUIRE(si.rows_b > 0 && si.col0 >= 0 && ld >= si.col0 + si.rows_b && (!sym || rows_a == si.rows_b),
                    "nsr_contract: bad segment %d (rows_b=%lld col0=%lld ld=%lld)", s, (long long)si.rows_b,
                    (long long)si.col0, (long long)ld);
        NSR_REQUIRE(si.mode != NSR_MODE_COEX || n_segs == 1, "nsr_contract: NSR_MODE_COEX (mirrored) takes one segment");
        NSR_REQUIRE(segs[s].slices && si.qb && ((uintptr_t)segs[s].slices & 15) == 0, "nsr_contract: segment %d: bad planes", s);
        NSR_REQUIRE(si.mode == NSR_MODE_RAW || si.vb != nullptr, "nsr_contract: segment %d: var missing", s);
        need_p = need_p || si.mode != NSR_MODE_RAW;
    }
    NSR_REQUIRE(!need_p || (P != nullptr && var_a != nullptr && dof_a > 0.0), "nsr_contract: P / var / dof missing");
    if (n_tiles == 0) return 0;
    NSR_REQUIRE(host_tiles != nullptr && n_tiles > 0 && n_tiles < (1ll << 30), "nsr_contract: bad tile list");
    const int64_t tr_max = (rows_a + NSR_TILE - 1) / NSR_TILE;
    std::vector<int32_t> packed((size_t)n_tiles * 2);
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int32_t sg = tile_stride == 3 ? host_tiles[3 * t] : 0;
        const int32_t tr = host_tiles[tile_stride * t + tile_stride - 2], tc = host_tiles[tile_stride * t + tile_stride - 1];
        NSR_REQUIRE(sg >= 0 && sg < n_segs, "nsr_contract: tile %lld: segment %d out of range", (long long)t, sg);
        const int64_t tc_max = (segs[sg].info.rows_b + NSR_TILE - 1) / NSR_TILE;
        NSR_REQUIRE(tr >= 0 && tr < tr_max && tc >= 0 && tc < tc_max && tc < (1 << 24),
                    "nsr_contract: tile %lld = (%d,%d) out of range", (long long)t, tr, tc);
        const int m = segs[sg].info.mode;
        NSR_REQUIRE(!(m == NSR_MODE_COEX || m == NSR_MODE_COEX_UPPER) || tr <= tc, "nsr_contract: COEX tiles must have row <= col");
        packed[2 * t] = tr;
        packed[2 * t + 1] = tc | (sg << 24);
    }
    cudaStream_t st = (cudaStream_t)stream;
    NSR_CHECK(cudaSetDevice(ctx->device));
    // the cta_group::2 kernel works on 256 x 128 pair tiles: fold (tr, tc) into (tr/2, tc, half mask),
    // keeping first-appearance order (the caller's order carries the L2-locality plan)
    const bool pair = engine == NSR_ENGINE_UMMA && nsr_umma_pair != 0 && n_segs == 1 && n_slices_a == n_slices_b &&
                      n_slices_a >= 3;
    std::vector<int32_t> folded;
    int64_t n_upload = n_tiles * 2;
    const int32_t* upload = packed.data();
    int64_t n_entries = n_tiles;
    if (pair) {
        std::unordered_map<uint64_t, int64_t> seen;
        seen.reserve((size_t)n_tiles * 2);
        folded.reserve((size_t)n_tiles * 3);
        for (int64_t t = 0; t < n_tiles; ++t) {
            const int32_t tr = packed[2 * t], tc = packed[2 * t + 1];
            const uint64_t key = ((uint64_t)(uint32_t)(tr >> 1) << 32) | (uint32_t)tc;
            auto it = seen.find(key);
            if (it == seen.end()) {
                seen.emplace(key, (int64_t)folded.size() / 3);
                folded.push_back(tr >> 1);
                folded.push_back(tc);
                folded.push_back(1 << (tr & 1));
            } else {
                folded[(size_t)it->second * 3 + 2] |= 1 << (tr & 1);
            }
        }
        n_entries = (int64_t)folded.size() / 3;
        n_upload = n_entries * 3;
        upload = folded.data();
    }
    // The tile list travels through a pinned staging buffer owned by the context (a pageable source
    // would make cudaMemcpyAsync block the host until the stream drains, and would tie the lifetime of
    // the vectors above to that staging behaviour); an event guards its reuse by the next call.
    if ((size_t)n_upload > ctx->tiles_cap) {
        NSR_CHECK(cudaStreamSynchronize(st));
        if (ctx->tiles_dev) NSR_CHECK(cudaFree(ctx->tiles_dev));
        if (ctx->tiles_pinned) NSR_CHECK(cudaFreeHost(ctx->tiles_pinned));
        ctx->tiles_dev = nullptr;
        ctx->tiles_pinned = nullptr;
        ctx->tiles_cap = 0;
        NSR_CHECK(cudaMalloc(&ctx->tiles_dev, (size_t)n_upload * sizeof(int32_t) * 2));
        NSR_CHECK(cudaMallocHost(&ctx->tiles_pinned, (size_t)n_upload * sizeof(int32_t) * 2));
        ctx->tiles_cap = (size_t)n_upload * 2;
    }
    if (ctx->tiles_event == nullptr) NSR_CHECK(cudaEventCreateWithFlags(&ctx->tiles_event, cudaEventDisableTiming));
    else NSR_CHECK(cudaEventSynchronize(ctx->tiles_event));       // previous upload has left the staging buffer
    memcpy(ctx->tiles_pinned, upload, (size_t)n_upload * sizeof(int32_t));
    // NOT a cudaMemcpyAsync: a host-to-device copy queues on the H2D copy engine, and in the streamed pipelines that
    // engine is busy with the next gigabyte-sized chunk of the expression matrix - the tile list then waits ~22 ms
    // behind it and so does the contraction (measured: every strip's launch started one chunk late, 350 -> 3xx ms
    // end to end at 100k x 20k).  A small kernel reads the page-locked list over PCIe instead.
    {
        const int64_t n_vec = (n_upload + 3) / 4;            // the buffers are allocated with twice the room asked for
        const int blocks = (int)((n_vec + 255) / 256 < 16 ? (n_vec + 255) / 256 : 16);
        tiles_upload_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(reinterpret_cast<int4*>(ctx->tiles_dev),
                                                                    reinterpret_cast<const int4*>(ctx->tiles_pinned), n_vec);
        NSR_CHECK(cudaGetLastError());
    }
    NSR_CHECK(cudaEventRecord(ctx->tiles_event, st));

    ContractParams ep;
    ep.mode = mode;
    ep.n_groups = wmax - 1 < n_slices_a + n_slices_b - 1 ? wmax - 1 : n_slices_a + n_slices_b - 1;
    ep.rows_a = rows_a; ep.rows_b = segs[0].info.rows_b; ep.ld = ld;
    ep.qa = quantum_a; ep.va = var_a; ep.qb = segs[0].info.qb; ep.vb = segs[0].info.vb;
    ep.P = P; ep.out2 = out2;
    ep.inv_n = 1.0 / (double)n;
    ep.acc_in = 0;
    ep.raw_out = 0;
    ep.refine_r2 = -1.0;
    ep.need = nullptr;
    for (int g = 0; g < 4; ++g) ep.group_scale[g] = (g < ep.n_groups) ? ldexp(1.0, 8 * (ep.n_groups - 1 - g)) : 0.0;
    ep.scale_all = ldexp(1.0, 8 * (n_slices_a + n_slices_b - 1 - ep.n_groups));
    ep.pv = nsr_pval_params(need_p ? dof_a : 1.0);

    // Cell chunking: int32 accumulators are exact only while no partial sum can overflow; the caller
    // bounds that (Cauchy-Schwarz on the digit-plane energies) and passes the chunk length.  Chunks
    // but the last leave the float64 running sum in out2; the last one adds it and finishes.
    NSR_REQUIRE(k_chunk >= 0 && k_chunk % NSR_KBLOCK == 0, "nsr_contract: k_chunk must be a multiple of %d", NSR_KBLOCK);
    const int64_t chunk = (k_chunk == 0 || k_chunk >= n_pad) ? n_pad : k_chunk;
    if (engine == NSR_ENGINE_UMMA && !pair && n_segs == 1 && chunk == n_pad && mode != NSR_MODE_RAW && n_slices_a == 3 &&
        n_slices_b == 3 && wmax == 5 && nsr_adaptive_min_cells > 0 && n >= nsr_adaptive_min_cells) {
        // ---- adaptive schedule: 6 products everywhere, 8 where a pair is extremely significant
        const size_t want = 3 * (size_t)n_tiles + 4;
        if (want > ctx->refine_cap) {
            if (ctx->refine_dev) NSR_CHECK(cudaFree(ctx->refine_dev));
            ctx->refine_dev = nullptr;
            ctx->refine_cap = 0;
            NSR_CHECK(cudaMalloc(&ctx->refine_dev, want * 2 * sizeof(int32_t)));
            ctx->refine_cap = want * 2;
        }
        int* need = ctx->refine_dev;
        int32_t* list = ctx->refine_dev + n_tiles;
        int* count = ctx->refine_dev + 3 * n_tiles;
        NSR_CHECK(cudaMemsetAsync(need, 0, (size_t)n_tiles * sizeof(int32_t), st));
        NSR_CHECK(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
        ContractParams ep1 = ep;
        ep1.n_groups = 3;                                             // wmax = 4
        for (int g = 0; g < 4; ++g) ep1.group_scale[g] = (g < 3) ? ldexp(1.0, 8 * (2 - g)) : 0.0;
        ep1.scale_all = ldexp(1.0, 8 * (2 * 3 - 4));
        ep1.refine_r2 = kRefineZ2 / (double)n;
        ep1.need = need;
        int rc = nsr_launch_contract_umma(ctx, st, a_slices, rows_a, rows_alloc_a, 3, segs, 1, n_pad, 3, 4, ctx->tiles_dev,
                                          n_tiles, ep1, 0, n_pad, nullptr);
        if (rc) return rc;
        refine_compact_kernel<<<(unsigned)((n_tiles + 255) / 256 < 64 ? (n_tiles + 255) / 256 : 64), 256, 0, st>>>(
            need, ctx->tiles_dev, (int)n_tiles, list, count);
        NSR_CHECK(cudaGetLastError());
        return nsr_launch_contract_umma(ctx, st, a_slices, rows_a, rows_alloc_a, 3, segs, 1, n_pad, 3, 5, list, n_tiles, ep,
                                        0, n_pad, count);
    }
    // Few tiles, many cells (de: a few hundred groupings against every gene; the groupings' own Gram matrix):
    // with one CTA per tile the last wave of tiles leaves most SMs idle.  Split the cells into parts instead -
    // work items (tile, part) over all SMs, partial sums into per-part slabs, one small kernel that adds the
    // slabs in a fixed order and finishes.  The parts also serve as the overflow chunks (each part <= chunk).
    // Only the rectangular modes: their sums (two different variables, or an exact small-integer operand) stay
    // below 2^53, so the parts add up exactly and the result is the single-pass one bit for bit; co-expression
    // keeps its one-pass tiles (self-products of 24-bit values over 1e5 cells exceed 2^53, and its results are
    // promised to be identical across tilings and GPU counts).
    // Worth it when the one-CTA-per-tile launch would leave more than a quarter of the machine idle (its last wave):
    // the second kernel and the slabs cost about as much as a 20 % imbalance on these short launches (measured at
    // 300 x 10k x 50k cells: 237 tiles = 1.6 waves ran 0.52 ms one-pass, 0.6 ms split; 160 tiles over 1M cells ran
    // 10.6 ms one-pass, 7.9 ms split).
    const int64_t waves = (n_tiles + ctx->sm_count - 1) / ctx->sm_count;
    const bool unbalanced = 4 * n_tiles < 3 * waves * (int64_t)ctx->sm_count;
    if (engine == NSR_ENGINE_UMMA && !pair && n_segs == 1 && nsr_split_k != 0 && n_tiles < 2 * (int64_t)ctx->sm_count &&
        unbalanced && (mode == NSR_MODE_DE || mode == NSR_MODE_RAW)) {
        const int64_t target = 4 * (int64_t)ctx->sm_count;                  // ~4 waves of work items
        int64_t want = (target + n_tiles - 1) / n_tiles;
        const int64_t for_overflow = (n_pad + chunk - 1) / chunk;
        if (want < for_overflow) want = for_overflow;
        const int64_t max_parts = n_pad / (8 * NSR_KBLOCK) > 0 ? n_pad / (8 * NSR_KBLOCK) : 1;   // >= 1024 cells per part
        if (want > max_parts && max_parts >= for_overflow) want = max_parts;
        const int64_t slab = rows_a * ld;                                   // one (rows_a x ld) image of out2 per part
        const int parts = nsr_umma_parts(n_pad, n_slices_a, n_slices_b, (int)want);
        if (parts > 1 && parts >= for_overflow && (size_t)parts * (size_t)slab * sizeof(double) <= ((size_t)1 << 30)) {
            void* scratch = nullptr;
            if (nsr_scratch(ctx, (size_t)parts * (size_t)slab * sizeof(double), &scratch)) return 1;
            int rc = nsr_launch_contract_umma(ctx, st, a_slices, rows_a, rows_alloc_a, n_slices_a, segs, 1, n_pad, n_slices_b,
                                              wmax, ctx->tiles_dev, n_tiles, ep, 0, n_pad, nullptr, parts, (double*)scratch,
                                              slab);
            if (rc) return rc;
            return nsr_launch_contract_finish(st, ctx->tiles_dev, n_tiles, ep, segs[0].info, parts, (const double*)scratch,
                                              slab);
        }
    }
    for (int64_t c0 = 0; c0 < n_pad; c0 += chunk) {
        const int64_t c1 = c0 + chunk < n_pad ? c0 + chunk : n_pad;
        ep.acc_in = c0 > 0;
        ep.raw_out = c1 < n_pad;
        int rc;
        if (engine == NSR_ENGINE_SIMT) {
            rc = nsr_launch_contract_simt(st, a_slices, rows_alloc_a, n_slices_a, segs[0].slices, segs[0].rows_alloc,
                                          n_slices_b, n_pad, wmax, ctx->tiles_dev, n_tiles, ep, c0, c1);
            if (rc) {
                nsr_set_error("nsr_contract: SIMT launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        } else {
            rc = nsr_launch_contract_umma(ctx, st, a_slices, rows_a, rows_alloc_a, n_slices_a, segs, n_segs, n_pad,
                                          n_slices_b, wmax, ctx->tiles_dev, pair ? -n_entries : n_tiles, ep, c0, c1, nullptr);
            if (rc) return rc;
        }
    }
    return 0;
}

extern "C" int nsr_contract_ab(nsr_ctx* ctx, uintptr_t stream, int engine, int mode,
                               const int8_t* a_slices, int64_t rows_a, int64_t rows_alloc_a, int n_slices_a,
                               const double* quantum_a, const double* var_a, const int8_t* b_slices,
                               int64_t rows_b, int64_t rows_alloc_b, int n_slices_b, const double* quantum_b,
                               const double* var_b, int64_t n, int64_t n_pad, int n_products,
                               const int32_t* host_tiles, int64_t n_tiles, double dof_a,
                               double* P, double* out2, int64_t ld, int64_t k_chunk) {
    NsrSegOperand sg;
    sg.slices = b_slices;
    sg.rows_alloc = rows_alloc_b;
    sg.info.qb = quantum_b; sg.info.vb = var_b; sg.info.rows_b = rows_b; sg.info.col0 = 0; sg.info.mode = mode;
    sg.info.ready = nullptr; sg.info.ready_value = 0; sg.info.done = nullptr;
    const bool mir = mode == NSR_MODE_COEX;          // both triangles of the same matrix
    sg.info.mP = mir ? P : nullptr; sg.info.mO = mir ? out2 : nullptr; sg.info.ldm = ld;
    return contract_impl(ctx, stream, engine, mode, a_slices, rows_a, rows_alloc_a, n_slices_a, quantum_a, var_a, &sg, 1,
                         n_slices_b, n, n_pad, n_products, host_tiles, 2, n_tiles, dof_a, P, out2, ld, k_chunk);
}

extern "C" int nsr_contract(nsr_ctx* ctx, uintptr_t stream, int engine, int mode,
                            const int8_t* a_slices, int64_t rows_a, int64_t rows_alloc_a,
                            const double* quantum_a, const double* var_a, const int8_t* b_slices,
                            int64_t rows_b, int64_t rows_alloc_b, const double* quantum_b,
                            const double* var_b, int64_t n, int64_t n_pad, int n_slices,
                            int n_products, const int32_t* host_tiles, int64_t n_tiles, double dof_a,
                            double* P, double* out2, int64_t ld, int64_t k_chunk) {
    return nsr_contract_ab(ctx, stream, engine, mode, a_slices, rows_a, rows_alloc_a, n_slices, quantum_a, var_a, b_slices,
                           rows_b, rows_alloc_b, n_slices, quantum_b, var_b, n, n_pad, n_products, host_tiles, n_tiles, dof_a,
                           P, out2, ld, k_chunk);
}

extern "C" int nsr_contract_segments(nsr_ctx* ctx, uintptr_t stream, const int8_t* a_slices, int64_t rows_a,
                                     int64_t rows_alloc_a, const double* quantum_a, const double* var_a,
                                     int64_t n, int64_t n_pad, int n_slices, int n_products,
                                     const nsr_segment* segments, int n_segments, const int32_t* host_tiles,
                                     int64_t n_tiles, double dof_a, double* P, double* out2, int64_t ld,
                                     int64_t k_chunk) {
    NSR_REQUIRE(segments != nullptr && n_segments >= 1 && n_segments <= NSR_MAX_SEGMENTS, "nsr_contract_segments: %d segments (1..%d)",
                n_segments, NSR_MAX_SEGMENTS);
    NsrSegOperand sg[NSR_MAX_SEGMENTS];
    for (int s = 0; s < n_segments; ++s) {
        const nsr_segment& in = segments[s];
        sg[s].slices = in.b_slices;
        sg[s].rows_alloc = in.rows_alloc_b;
        sg[s].info.qb = in.quantum_b; sg[s].info.vb = in.var_b; sg[s].info.rows_b = in.rows_b; sg[s].info.col0 = in.col0;
        sg[s].info.mode = in.diagonal ? NSR_MODE_COEX_UPPER : NSR_MODE_COEX_RECT;
        sg[s].info.ready = in.ready; sg[s].info.ready_value = in.ready_value; sg[s].info.done = in.done;
        NSR_REQUIRE((in.mirror_P == nullptr) == (in.mirror_out2 == nullptr) && (in.mirror_P == nullptr || in.ld_mirror >= rows_a),
                    "nsr_contract_segments: segment %d: bad mirror (ld_mirror=%lld)", s, (long long)in.ld_mirror);
        sg[s].info.mP = in.mirror_P; sg[s].info.mO = in.mirror_out2; sg[s].info.ldm = in.ld_mirror;
    }
    return contract_impl(ctx, stream, NSR_ENGINE_UMMA, NSR_MODE_COEX_RECT, a_slices, rows_a, rows_alloc_a, n_slices, quantum_a,
                         var_a, sg, n_segments, n_slices, n, n_pad, n_products, host_tiles, 3, n_tiles, dof_a, P, out2, ld,
                         k_chunk);
}

// Stream-ordered 32-bit flag operations (cuStreamWriteValue32 / cuStreamWaitValue32): no kernel, no SM.
typedef CUresult (*StreamValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static int stream_value_op(nsr_ctx* ctx, const char* sym, uintptr_t stream, uint32_t* flag, uint32_t value, unsigned flags) {
    NSR_REQUIRE(ctx != nullptr && flag != nullptr, "%s: null argument", sym);
    NSR_CHECK(cudaSetDevice(ctx->device));
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    cudaError_t e = cudaGetDriverEntryPoint(sym, &fn, cudaEnableDefault, &qres);
    NSR_REQUIRE(e == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn != nullptr, "%s unavailable", sym);
    CUresult r = ((StreamValueFn)fn)((CUstream)stream, (CUdeviceptr)(uintptr_t)flag, value, flags);
    NSR_REQUIRE(r == CUDA_SUCCESS, "%s failed with CUresult %d", sym, (int)r);
    return 0;
}
extern "C" int nsr_stream_signal(nsr_ctx* ctx, uintptr_t stream, uint32_t* flag, uint32_t value) {
    return stream_value_op(ctx, "cuStreamWriteValue32", stream, flag, value, 0 /* CU_STREAM_WRITE_VALUE_DEFAULT */);
}
extern "C" int nsr_stream_wait_geq(nsr_ctx* ctx, uintptr_t stream, uint32_t* flag, uint32_t value) {
    return stream_value_op(ctx, "cuStreamWaitValue32", stream, flag, value, 0 /* CU_STREAM_WAIT_VALUE_GEQ */);
}

extern "C" int nsr_last_refined(nsr_ctx* ctx, uintptr_t stream, int64_t n_tiles, int64_t* refined) {
    NSR_REQUIRE(ctx && refined && n_tiles >= 1, "nsr_last_refined: bad arguments");
    NSR_REQUIRE(ctx->refine_dev != nullptr && 3 * (size_t)n_tiles + 1 <= ctx->refine_cap, "nsr_last_refined: no adaptive launch of that size yet");
    NSR_CHECK(cudaSetDevice(ctx->device));
    int32_t c = 0;
    NSR_CHECK(cudaMemcpyAsync(&c, ctx->refine_dev + 3 * n_tiles, sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    NSR_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    *refined = c;
    return 0;
}

extern "C" int nsr_copy2d(nsr_ctx* ctx, uintptr_t stream, void* dst, int64_t dst_pitch, const void* src,
                          int64_t src_pitch, int64_t width_bytes, int64_t height, int kind) {
    NSR_REQUIRE(ctx != nullptr && dst && src && width_bytes >= 0 && height >= 0 && dst_pitch >= width_bytes &&
                    src_pitch >= width_bytes && (kind == 0 || kind == 1),
                "nsr_copy2d: bad arguments");
    if (width_bytes == 0 || height == 0) return 0;
    NSR_CHECK(cudaSetDevice(ctx->device));
    NSR_CHECK(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes, (size_t)height,
                                kind == 0 ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return 0;
}

extern "C" int nsr_copy_peer(nsr_ctx* ctx, uintptr_t stream, void* dst, const void* src, int src_device, int64_t nbytes) {
    NSR_REQUIRE(ctx != nullptr && dst && src && nbytes >= 0 && src_device >= 0, "nsr_copy_peer: bad arguments");
    if (nbytes == 0) return 0;
    NSR_CHECK(cudaSetDevice(ctx->device));
    if (src_device != ctx->device) {
        int can = 0;
        NSR_CHECK(cudaDeviceCanAccessPeer(&can, ctx->device, src_device));
        if (can) {                                 // direct NVLink / PCIe peer path; otherwise the driver stages through the host
            cudaError_t e = cudaDeviceEnablePeerAccess(src_device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else NSR_CHECK(e);
        }
    }
    NSR_CHECK(cudaMemcpyPeerAsync(dst, ctx->device, src, src_device, (size_t)nbytes, (cudaStream_t)stream));
    return 0;
}

namespace {
__global__ void pvalue_kernel(const double* __restrict__ r2, const double* __restrict__ a, int64_t row_len,
                              int64_t count, double* __restrict__ P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const NsrPvalParams p = nsr_pval_params(a[i / row_len]);
    P[i] = nsr_pvalue_r2(r2[i], p);
}
}  // namespace

extern "C" int nsr_pvalue(nsr_ctx* ctx, uintptr_t stream, const double* r2, const double* a, int64_t row_len,
                          int64_t count, double* P) {
    NSR_REQUIRE(ctx != nullptr && r2 && a && P && row_len > 0 && count >= 0, "nsr_pvalue: bad arguments");
    if (count == 0) return 0;
    NSR_CHECK(cudaSetDevice(ctx->device));
    pvalue_kernel<<<(unsigned)((count + 127) / 128), 128, 0, (cudaStream_t)stream>>>(r2, a, row_len, count, P);
    NSR_CHECK(cudaGetLastError());
    return 0;
}
