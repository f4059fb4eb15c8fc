// tcgen05 / TMEM / TMA contraction over cells for sm_100a: the product path.
//
// For an output tile (128 rows of A) x (128 rows of B) the kernel forms, for every kept pair
// of digit planes (a, b), the exact integer sum over cells  sum_k dA_a[i,k] dB_b[j,k]  with
// tcgen05.mma.kind::i8 (int8 x int8 -> int32 in TMEM; integer accumulation is exact, which a
// float accumulator - round-toward-zero in the tensor pipe - is not).  Pairs with equal a + b
// share one 128-column TMEM accumulator.  The epilogue warps read the accumulators with
// tcgen05.ld, combine them in float64, scale by the row quanta and evaluate r^2 and the exact
// P-value in registers, so P and dot are written once (reference association.py:234-249 and
// the assembly at :1036-1057).
//
// Roles (320 threads, persistent over a tile list, one CTA per SM):
//   warp 0 lane 0  TMA producer: 2S boxes (128 rows x KB bytes, swizzled) per k-block
//   warp 1         TMEM allocator; lane 0 issues the MMAs and commits to mbarriers
//   warps 2..9     epilogue: warp w reads TMEM lane quadrant w%4, columns 64*((w-2)/4)..+64.
// The P-value arithmetic is latency-bound float64 (~330 us per tile with 4 epilogue warps, as long
// as 40 % of a 100k-cell tile), hence 8 epilogue warps, two per scheduler.
//
// This file is compiled twice.  contract_umma.cu itself is the one-pass kernel and everything on the host side;
// contract_umma_splitk.cu defines NSR_SPLITK_TU and includes it: the same kernel with work items (tile, part
// of the cells) - see UmmaArgs::n_parts - exported as nsr_launch_contract_umma_splitk.  Two translation units
// rather than a template flag so that the one-pass kernel's argument block, code and registers are exactly
// what they are without the split (power-bound at the headline size: profiles/r02_splitk_ab.md).
#include <cuda.h>
#ifndef NSR_SPLITK_TU
#define NSR_SPLITK_TU 0
#endif

#include "epilogue.cuh"

extern int nsr_epi_warps;
extern int nsr_umma_stack;
extern int nsr_umma_dynamic;

namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr uint32_t kTmemCols = 512;
constexpr int kSmemBudget = 200 * 1024;       // operand ring; barriers live in static smem

constexpr int kMaxSegs = NSR_MAX_SEGMENTS;

// tensor maps of the segments' B operands (a __grid_constant__ kernel parameter: TMA descriptors must
// live in param / const / global space)
struct SegMaps {
    CUtensorMap b[kMaxSegs];
};

struct UmmaArgs {
    const int32_t* tiles;         // pairs (tile_row, tile_col | segment << 24)
    int n_tiles;
    int num_kb;                   // k-blocks of KB cells in this launch
    int kb_begin;                 // first k-block (cell chunking)
    int stack_b;                  // 1: N = 256 MMAs over two stacked B planes (single-CTA kernel)
    int epi_sleep_ns;             // back-off of the epilogue warps while they wait for a tile
    int* tile_counter;            // dynamic tile scheduler (single-CTA kernel): next unclaimed list index,
                                  // zeroed before the launch; nullptr = static round-robin
    const int* n_tiles_dev;       // if set, the tile count is read from device memory (second phase of the
                                  // adaptive schedule: the list was compacted on the device)
    int n_segs;
#if NSR_SPLITK_TU
    // split over the cells: work item w = (tile w % n_tiles, part w / n_tiles); part p contracts k-blocks
    // [p * kb_per_part, (p + 1) * kb_per_part) and stores its unscaled partial sums in slab p of part_out
    int n_parts;
    int kb_per_part;
    double* part_out;
    int64_t part_stride;          // doubles per slab
#endif
    SegInfo seg[kMaxSegs];
    ContractParams ep;
};

constexpr int kTileRing = 8;      // tile-index hand-off ring; the producer leads the epilogue by <= 3 tiles

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must abort the launch, never hang the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {       // ~2 s
            printf("nsr umma: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// Waiting epilogue warps must not spin on the issue ports (they wait most of a tile's duration).
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, int sleep_ns) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (sleep_ns > 0) __nanosleep(sleep_ns);
        if (clock64() - t0 > 8000000000ll) {
            printf("nsr umma: epilogue mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// A segment's B operand arrives from another GPU (copy-engine pull into local memory) while this
// launch is already running: the producer waits for the stream-ordered flag write that follows the
// copy before its first TMA read of that segment.  Bounded like every wait in this file.
__device__ __forceinline__ void wait_ready(const uint32_t* flag, uint32_t want) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - want) >= 0) break;
        __nanosleep(200);
        if (clock64() - t0 > 20000000000ll) {      // ~10 s
            printf("nsr umma: segment flag timeout block %d (have %u, want %u)\n", blockIdx.x, v, want);
            __trap();
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");   // later async-proxy (TMA) reads see the copied planes
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
template <int N>
__device__ __forceinline__ void tc_ldN(uint32_t taddr, uint32_t (&v)[N]);
template <>
__device__ __forceinline__ void tc_ldN<16>(uint32_t taddr, uint32_t (&v)[16]) { tc_ld16(taddr, v); }
template <>
__device__ __forceinline__ void tc_ldN<8>(uint32_t taddr, uint32_t (&v)[8]) { tc_ld8(taddr, v); }
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, rows of KB bytes, swizzle width == KB (128B or 64B), 8-row groups dense
template <int KB>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    constexpr uint64_t layout = (KB == 128) ? 2ull : 4ull;       // SWIZZLE_128B / SWIZZLE_64B
    constexpr uint64_t sbo = (8ull * KB) >> 4;                   // next 8-row group
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) |
           (layout << 61);
}
// int8 x int8 -> int32, A and B K-major, M = 128, N = 128
constexpr uint32_t kInstrDesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);


__device__ __forceinline__ void mbar_arrive_cluster_addr(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// One 128 x 128 sub-tile for the 8 epilogue warps of a CTA: warp w owns TMEM lane quadrant w%4
// (a hardware rule) and columns 64*((w-2)/4) .. +64, 16 at a time.  empty_bar: the barrier the
// MMA issuer waits on before reusing TMEM (shared::cta address, or shared::cluster if `remote`).
//
// Measured alternatives (profiles/r01_epilogue_ab.md, 100k x 20k, 12 steps under the 1 kW cap):
// draining all accumulators into registers first and releasing TMEM early (so the next tile's
// MMAs overlap the P-value arithmetic) is 13 % faster in a short burst but 25 % SLOWER sustained:
// the chip is power-bound here, and that variant costs more energy per tile than it saves time.
template <int GROUPS, int EW>
__device__ __forceinline__ void epilogue_tile(const ContractParams& ep, const SegInfo& sg, uint32_t tmem_base, int warp,
                                              int lane, int tr, int tc, bool wanted, uint32_t empty_bar, bool remote,
#if NSR_SPLITK_TU
                                              int tile_idx, double* part) {
#else
                                              int tile_idx = -1) {
#endif
    constexpr int kCols = NSR_TILE / (EW / 4);          // columns per epilogue warp
    constexpr int CH = EW > 8 ? 8 : 16;                 // columns per TMEM read (register budget)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                   // column group of this warp
    const int64_t i = (int64_t)tr * NSR_TILE + quad * 32 + lane;
    const bool row_ok = wanted && i < ep.rows_a;
    const double qi = row_ok ? ep.qa[i] : 0.0;
    const double vi = (row_ok && ep.va) ? ep.va[i] : 1.0;
    // the transposed copy: none for a tile on the diagonal of a symmetric segment (it holds both (i, j) and (j, i))
    double* const mP = (sg.mP != nullptr && (tr != tc || sg.mode == NSR_MODE_COEX_RECT)) ? sg.mP : nullptr;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    bool refine = false;
    if (wanted) {
#pragma unroll 1
        for (int c0 = half * kCols; c0 < half * kCols + kCols; c0 += CH) {
            uint32_t v[GROUPS][CH];
#pragma unroll
            for (int grp = 0; grp < GROUPS; ++grp) tc_ldN<CH>(lane_base + grp * NSR_TILE + c0, v[grp]);
            tc_ld_wait();
            const int64_t j0 = (int64_t)tc * NSR_TILE + c0;
            if (row_ok && j0 < sg.rows_b) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int64_t j = j0 + c;
                    if (j < sg.rows_b) {
                        int32_t a4[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int grp = 0; grp < GROUPS; ++grp) a4[grp] = (int32_t)v[grp][c];
                        // (the per-row scales of a remote segment are written by a copy engine during this
                        // launch, but before the segment's flag: no SM has them in L1 earlier, and L1 does not
                        // survive a launch boundary - plain cached loads are safe)
#if NSR_SPLITK_TU
                        // split over the cells: the unscaled integer sum of this part (contract_finish_kernel adds them)
                        part[i * ep.ld + sg.col0 + j] = nsr_combine(ep, a4);
#else
                        refine |= nsr_finish(ep, sg.mode, sg.col0, i, j, qi, vi, sg.qb[j], sg.vb ? sg.vb[j] : 1.0,
                                             nsr_combine(ep, a4), mP, sg.mO, sg.ldm);
#endif
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
        if (remote) mbar_arrive_cluster_addr(empty_bar);
        else mbar_arrive(empty_bar);
        if (sg.done != nullptr) {            // this warp's part of the tile is in memory: count it (D2H streams wait on it)
            __threadfence_system();
            atomicAdd(sg.done, 1u);
        }
    }
    // adaptive schedule, first phase: any element beyond the threshold sends the tile to the second phase
    if (ep.need != nullptr && tile_idx >= 0 && __any_sync(0xffffffffu, refine) && lane == 0) atomicOr(&ep.need[tile_idx], 1);
}

// int8 x int8 -> int32, M = 128, N = 256 (two stacked B planes)
constexpr uint32_t kInstrDescN256 = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

// MMA schedule with stacked B planes: entries (a, b, wide, first): A plane a times B plane b
// (wide: planes b and b+1 as one N = 256 operand); `first` = the entry initialises its group(s).
template <int SA, int SB, int WMAX> struct Sched;
template <> struct Sched<3, 3, 5> {      // products (0,0)(0,1) | (0,2) | (1,0)(1,1) | (1,2) | (2,0)(2,1)
    static constexpr int kCount = 5;
    int a[5] = {0, 0, 1, 1, 2}, b[5] = {0, 2, 0, 2, 0};
    bool wide[5] = {true, false, true, false, true}, first[5] = {true, true, false, true, false};
};
template <> struct Sched<3, 3, 4> {      // (0,0)(0,1) | (0,2) | (1,0)(1,1) | (2,0)
    static constexpr int kCount = 4;
    int a[4] = {0, 0, 1, 2}, b[4] = {0, 2, 0, 0};
    bool wide[4] = {true, false, true, false}, first[4] = {true, true, false, false};
};
template <> struct Sched<4, 4, 5> {      // (0,0)(0,1) | (0,2)(0,3) | (1,0)(1,1) | (1,2) | (2,0)(2,1) | (3,0)
    static constexpr int kCount = 6;
    int a[6] = {0, 0, 1, 1, 2, 3}, b[6] = {0, 2, 0, 2, 0, 0};
    bool wide[6] = {true, true, true, false, true, false}, first[6] = {true, true, false, false, false, false};
};
// single-plane A operand (exact small integers: binary groupings after the Hadamard mix, see
// nsr_residualize_exact): every product of the B planes with the one A plane
template <> struct Sched<1, 3, 4> {      // (0,0)(0,1) | (0,2)
    static constexpr int kCount = 2;
    int a[2] = {0, 0}, b[2] = {0, 2};
    bool wide[2] = {true, false}, first[2] = {true, true};
};
template <> struct Sched<1, 4, 5> {      // (0,0)(0,1) | (0,2)(0,3)
    static constexpr int kCount = 2;
    int a[2] = {0, 0}, b[2] = {0, 2};
    bool wide[2] = {true, true}, first[2] = {true, true};
};
template <> struct Sched<1, 1, 2> {      // (0,0)
    static constexpr int kCount = 1;
    int a[1] = {0}, b[1] = {0};
    bool wide[1] = {false}, first[1] = {true};
};

template <int SA, int SB, int WMAX, int KB>
struct Cfg {
    static constexpr int kSliceBytes = NSR_TILE * KB;
    static constexpr int kStageBytes = (SA + SB) * kSliceBytes;
    static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kGroups = (WMAX - 1) < (SA + SB - 1) ? (WMAX - 1) : (SA + SB - 1);
    static_assert(kStages >= 1, "stage does not fit");
    static_assert(kGroups * NSR_TILE <= (int)kTmemCols, "TMEM overflow");
};

template <int SA, int SB, int WMAX, int KB, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
contract_umma_kernel(const __grid_constant__ CUtensorMap map_a,
                     const __grid_constant__ SegMaps maps, const __grid_constant__ UmmaArgs g) {
    using C = Cfg<SA, SB, WMAX, KB>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_tmem_full, bar_tmem_empty, bar_tile[kTileRing];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int s_tile[kTileRing];

    uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t ring_u32 = smem_u32(ring);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int s = 0; s < kTileRing; ++s) mbar_init(smem_u32(&bar_tile[s]), 1);
        mbar_init(smem_u32(&bar_tmem_full), 1);
        mbar_init(smem_u32(&bar_tmem_empty), EW);         // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        // Tiles are claimed from a global counter (CTAs that start late or run on a slower SM simply
        // take fewer) and handed to the MMA and epilogue warps through a small shared-memory ring.
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t seg_seen = 0;                        // segments whose ready flag this CTA has observed
            const int n_tiles = g.n_tiles_dev ? *g.n_tiles_dev : g.n_tiles;
#if NSR_SPLITK_TU
            const int n_items = n_tiles * g.n_parts;
#else
            const int n_items = n_tiles;
#endif
            for (int it = 0;; ++it) {
                int w = g.tile_counter ? atomicAdd(g.tile_counter, 1) : (int)(blockIdx.x + it * gridDim.x);
                if (w >= n_items) w = -1;
                s_tile[it % kTileRing] = w;
                mbar_arrive(smem_u32(&bar_tile[it % kTileRing]));
                if (w < 0) break;
#if NSR_SPLITK_TU
                const int t = w % n_tiles;
                const int kb0 = (w / n_tiles) * g.kb_per_part, kb1 = min(g.num_kb, kb0 + g.kb_per_part);
#else
                const int t = w;
                constexpr int kb0 = 0;
                const int kb1 = g.num_kb;
#endif
                const int tcs = g.tiles[2 * t + 1];
                const int sidx = tcs >> 24;
                const int row_a = g.tiles[2 * t] * NSR_TILE, row_b = (tcs & 0xFFFFFF) * NSR_TILE;
                if (g.seg[sidx].ready != nullptr && !((seg_seen >> sidx) & 1u)) {
                    wait_ready(g.seg[sidx].ready, g.seg[sidx].ready_value);
                    seg_seen |= 1u << sidx;
                }
                const CUtensorMap* mb = &maps.b[sidx];
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
                    const uint32_t full = smem_u32(&bar_full[stage]);
                    mbar_expect_tx(full, C::kStageBytes);
                    const uint32_t base = ring_u32 + stage * C::kStageBytes;
#pragma unroll
                    for (int s = 0; s < SA; ++s)
                        tma_load_3d(&map_a, full, base + s * C::kSliceBytes, (g.kb_begin + kb) * KB, row_a, s);
#pragma unroll
                    for (int s = 0; s < SB; ++s)
                        tma_load_3d(mb, full, base + (SA + s) * C::kSliceBytes, (g.kb_begin + kb) * KB, row_b, s);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int it = 0;; ++it) {
                mbar_wait(smem_u32(&bar_tile[it % kTileRing]), (it / kTileRing) & 1);
                const int w = s_tile[it % kTileRing];
                if (w < 0) break;
#if NSR_SPLITK_TU
                const int kb0 = (w / g.n_tiles) * g.kb_per_part, kb1 = min(g.num_kb, kb0 + g.kb_per_part);
#else
                constexpr int kb0 = 0;
                const int kb1 = g.num_kb;
#endif
                mbar_wait(smem_u32(&bar_tmem_empty), tphase ^ 1);
                tc_fence_after();
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&bar_full[stage]), phase);
                    tc_fence_after();
                    const uint32_t base = ring_u32 + stage * C::kStageBytes;
#pragma unroll
                    for (int ks = 0; ks < KB / 32; ++ks) {
                        if (g.stack_b) {
                            // Two adjacent B planes are contiguous in the stage, i.e. one 256-row
                            // K-major tile: A(a) x [B(b); B(b+1)] with N = 256 accumulates product
                            // (a,b) into group a+b and (a,b+1) into group a+b+1 (adjacent TMEM
                            // columns) while reading the A tile once: 52 KB instead of 64 KB of
                            // operands per k-step, and 5 instructions instead of 8 (S=3, 8 products).
#pragma unroll
                            for (int e = 0; e < Sched<SA, SB, WMAX>::kCount; ++e) {
                                constexpr Sched<SA, SB, WMAX> sc{};
                                const int a = sc.a[e], b = sc.b[e];
                                const uint64_t da = make_smem_desc<KB>(base + a * C::kSliceBytes) + (uint64_t)(2 * ks);
                                const uint64_t db = make_smem_desc<KB>(base + (SA + b) * C::kSliceBytes) + (uint64_t)(2 * ks);
                                const uint32_t acc = (kb > kb0 || ks > 0 || !sc.first[e]) ? 1u : 0u;
                                tc_mma_i8(tmem_base + (a + b) * NSR_TILE, da, db, sc.wide[e] ? kInstrDescN256 : kInstrDesc, acc);
                            }
                        } else {
#pragma unroll
                            for (int a = 0; a < SA; ++a) {
#pragma unroll
                                for (int b = 0; b < SB; ++b) {
                                    if (a + b + 2 <= WMAX) {
                                        const int grp = a + b;
                                        // first product of its group in this (a asc, b asc) order
                                        const bool first = (a == (grp > SB - 1 ? grp - (SB - 1) : 0));
                                        const uint64_t da = make_smem_desc<KB>(base + a * C::kSliceBytes) + (uint64_t)(2 * ks);
                                        const uint64_t db = make_smem_desc<KB>(base + (SA + b) * C::kSliceBytes) + (uint64_t)(2 * ks);
                                        const uint32_t acc = (kb > kb0 || ks > 0 || !first) ? 1u : 0u;
                                        tc_mma_i8(tmem_base + grp * NSR_TILE, da, db, kInstrDesc, acc);
                                    }
                                }
                            }
                        }
                    }
                    tc_commit(smem_u32(&bar_empty[stage]));      // frees the stage when MMAs retire
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit(smem_u32(&bar_tmem_full));             // accumulators complete
                tphase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9)
        uint32_t tphase = 0;
        for (int it = 0;; ++it) {
            mbar_wait_backoff(smem_u32(&bar_tile[it % kTileRing]), (it / kTileRing) & 1, g.epi_sleep_ns);
            const int w = s_tile[it % kTileRing];
            if (w < 0) break;
#if NSR_SPLITK_TU
            const int t = w % g.n_tiles;
#else
            const int t = w;
#endif
            const int tr = g.tiles[2 * t], tcs = g.tiles[2 * t + 1];
            mbar_wait_backoff(smem_u32(&bar_tmem_full), tphase, g.epi_sleep_ns);
            tc_fence_after();
#if NSR_SPLITK_TU
            epilogue_tile<C::kGroups, EW>(g.ep, g.seg[tcs >> 24], tmem_base, warp, lane, tr, tcs & 0xFFFFFF, true,
                                          smem_u32(&bar_tmem_empty), false, t,
                                          g.part_out + (int64_t)(w / g.n_tiles) * g.part_stride);
#else
            epilogue_tile<C::kGroups, EW>(g.ep, g.seg[tcs >> 24], tmem_base, warp, lane, tr, tcs & 0xFFFFFF, true,
                                          smem_u32(&bar_tmem_empty), false, t);
#endif
            tphase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                     : "memory");
    }
}


#if !NSR_SPLITK_TU
// =============================================================================================
// cta_group::2 variant: a CTA pair (one cluster) computes a 256 x 128 output tile.  CTA r of the
// pair stages its own 128 rows of A and half (64 rows) of B, so per MMA each SM reads 6 KB of
// operands from shared memory instead of 8 KB and fills 72 KB instead of 96 KB per k-block: the
// 1-CTA kernel is bound by exactly that traffic (128 B/clk operand reads + 64 B/clk TMA fill
// against the 128 B/clk the SM can move; ncu: tensor pipe 68 % active).  The leader CTA issues
// tcgen05.mma.cta_group::2; both CTAs' TMA loads complete on the leader's full barrier;
// tcgen05.commit multicasts to both CTAs' empty / tmem_full barriers.
// Tile list entries are (tile_row/2, tile_col, mask): mask bit r = half r is wanted.
// =============================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst,
                                                int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tc_mma_i8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// int8 x int8 -> int32, K-major, M = 256 (pair), N = 128
constexpr uint32_t kInstrDesc2 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);

template <int S, int WMAX>
struct Cfg2 {
    static constexpr int kSliceA = NSR_TILE * 128;          // 128 rows x 128 B
    static constexpr int kSliceB = (NSR_TILE / 2) * 128;    // 64 rows x 128 B
    static constexpr int kStageBytes = S * (kSliceA + kSliceB);
    static constexpr int kStages = (216 * 1024) / kStageBytes;
    static constexpr int kGroups = WMAX - 1;
    static_assert(kStages >= 2 && kStages <= 8, "bad stage count");
};

template <int S, int WMAX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
contract_umma2_kernel(const __grid_constant__ CUtensorMap map_a,
                      const __grid_constant__ CUtensorMap map_b, const __grid_constant__ UmmaArgs g) {
    using C = Cfg2<S, WMAX>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[8], bar_empty[8], bar_tmem_full, bar_tmem_empty;
    __shared__ uint32_t tmem_base_slot;

    uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t ring_u32 = smem_u32(ring);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        mbar_init(smem_u32(&bar_tmem_full), 1);
        mbar_init(smem_u32(&bar_tmem_empty), 2 * kEpiWarps);   // epilogue warps of both CTAs of the pair
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < g.n_tiles; t += n_clusters) {
                const int row_a = g.tiles[3 * t] * (2 * NSR_TILE) + (int)rank * NSR_TILE;
                const int row_b = g.tiles[3 * t + 1] * NSR_TILE + (int)rank * (NSR_TILE / 2);
                for (int kb = 0; kb < g.num_kb; ++kb) {
                    mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1);
                    const uint32_t full_local = smem_u32(&bar_full[stage]);
                    const uint32_t full_leader = map_to_cta(full_local, 0);
                    if (leader) mbar_expect_tx(full_local, 2 * C::kStageBytes);
                    const uint32_t base = ring_u32 + stage * C::kStageBytes;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        tma_load_3d_2sm(&map_a, full_leader, base + s * C::kSliceA, (g.kb_begin + kb) * 128, row_a, s);
                        tma_load_3d_2sm(&map_b, full_leader, base + S * C::kSliceA + s * C::kSliceB, (g.kb_begin + kb) * 128, row_b, s);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            for (int t = cluster_id; t < g.n_tiles; t += n_clusters) {
                mbar_wait(smem_u32(&bar_tmem_empty), tphase ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < g.num_kb; ++kb) {
                    mbar_wait(smem_u32(&bar_full[stage]), phase);
                    tc_fence_after();
                    const uint32_t base = ring_u32 + stage * C::kStageBytes;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                        for (int a = 0; a < S; ++a) {
#pragma unroll
                            for (int b = 0; b < S; ++b) {
                                if (a + b + 2 <= WMAX) {
                                    const int grp = a + b;
                                    const bool first = (a == (grp > S - 1 ? grp - (S - 1) : 0));
                                    const uint64_t da = make_smem_desc<128>(base + a * C::kSliceA) + (uint64_t)(2 * ks);
                                    const uint64_t db = make_smem_desc<128>(base + S * C::kSliceA + b * C::kSliceB) + (uint64_t)(2 * ks);
                                    const uint32_t acc = (kb > 0 || ks > 0 || !first) ? 1u : 0u;
                                    tc_mma_i8_2sm(tmem_base + grp * NSR_TILE, da, db, kInstrDesc2, acc);
                                }
                            }
                        }
                    }
                    tc_commit_2sm(smem_u32(&bar_empty[stage]));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                tc_commit_2sm(smem_u32(&bar_tmem_full));
                tphase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
        uint32_t tphase = 0;
        const uint32_t empty_leader = map_to_cta(smem_u32(&bar_tmem_empty), 0);
        for (int t = cluster_id; t < g.n_tiles; t += n_clusters) {
            const int tr = g.tiles[3 * t] * 2 + (int)rank, tc = g.tiles[3 * t + 1];
            const bool wanted = (g.tiles[3 * t + 2] >> rank) & 1;
            mbar_wait_backoff(smem_u32(&bar_tmem_full), tphase, g.epi_sleep_ns);
            tc_fence_after();
            epilogue_tile<C::kGroups, kEpiWarps>(g.ep, g.seg[0], tmem_base, warp, lane, tr, tc, wanted, empty_leader, true);
            tphase ^= 1;
        }
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                     : "memory");
    }
}

#endif  // !NSR_SPLITK_TU

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int make_map(nsr_ctx* ctx, CUtensorMap* map, const int8_t* base, int64_t rows, int64_t rows_alloc,
             int64_t n_pad, int n_slices, int kb, int box_rows = NSR_TILE) {
    EncodeTiledFn fn = (EncodeTiledFn)ctx->encode_tiled;
    NSR_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
    cuuint64_t dims[3] = {(cuuint64_t)n_pad, (cuuint64_t)rows, (cuuint64_t)n_slices};
    cuuint64_t strides[2] = {(cuuint64_t)n_pad, (cuuint64_t)rows_alloc * (cuuint64_t)n_pad};
    cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    kb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NSR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

template <int SA, int SB, int WMAX, int KB, int EW>
int launch_ew(nsr_ctx* ctx, cudaStream_t st, const CUtensorMap& ma, const SegMaps& mb, const UmmaArgs& g) {
    using C = Cfg<SA, SB, WMAX, KB>;
    const int smem = C::kStages * C::kStageBytes + 1024;
    auto kern = contract_umma_kernel<SA, SB, WMAX, KB, EW>;
    NSR_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
#if NSR_SPLITK_TU
    const int64_t items = (int64_t)g.n_tiles * g.n_parts;
    const int grid = (g.n_tiles_dev == nullptr && items < ctx->sm_count) ? (int)items : ctx->sm_count;
#else
    const int grid = (g.n_tiles_dev == nullptr && g.n_tiles < ctx->sm_count) ? g.n_tiles : ctx->sm_count;
#endif
    kern<<<grid, 64 + 32 * EW, smem, st>>>(ma, mb, g);
    NSR_CHECK(cudaGetLastError());
    return 0;
}

template <int SA, int SB, int WMAX, int KB>
int launch(nsr_ctx* ctx, cudaStream_t st, const CUtensorMap& ma, const SegMaps& mb, const UmmaArgs& g) {
    if (nsr_epi_warps == 16) return launch_ew<SA, SB, WMAX, KB, 16>(ctx, st, ma, mb, g);
    return launch_ew<SA, SB, WMAX, KB, 8>(ctx, st, ma, mb, g);
}

#if !NSR_SPLITK_TU
template <int S, int WMAX>
int launch2(nsr_ctx* ctx, cudaStream_t st, const CUtensorMap& ma, const CUtensorMap& mb, const UmmaArgs& g) {
    using C = Cfg2<S, WMAX>;
    const int smem = C::kStages * C::kStageBytes + 1024;
    auto kern = contract_umma2_kernel<S, WMAX>;
    NSR_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int pairs = ctx->sm_count / 2;
    if (g.n_tiles < pairs) pairs = g.n_tiles;
    kern<<<2 * pairs, kThreads, smem, st>>>(ma, mb, g);
    NSR_CHECK(cudaGetLastError());
    return 0;
}

// Split over the cells, second step: one CTA per output tile adds the parts' slabs in a fixed order and
// turns the sum into the stored statistics (the same nsr_finish_sum as the fused epilogue).
__global__ void __launch_bounds__(256)
contract_finish_kernel(const int32_t* __restrict__ tiles, ContractParams ep, SegInfo sg, int n_parts,
                       const double* __restrict__ part, int64_t part_stride) {
    const int tr = tiles[2 * blockIdx.x], tc = tiles[2 * blockIdx.x + 1] & 0xFFFFFF;
    double* const mP = (sg.mP != nullptr && (tr != tc || sg.mode == NSR_MODE_COEX_RECT)) ? sg.mP : nullptr;
    for (int idx = threadIdx.x; idx < NSR_TILE * NSR_TILE; idx += 256) {
        const int64_t i = (int64_t)tr * NSR_TILE + (idx >> 7), j = (int64_t)tc * NSR_TILE + (idx & 127);
        if (i >= ep.rows_a || j >= sg.rows_b) continue;
        const int64_t at = i * ep.ld + sg.col0 + j;
        double acc = 0.0;                        // integer-valued partial sums: exact additions below 2^53
        for (int p = 0; p < n_parts; ++p) acc += part[(int64_t)p * part_stride + at];
        nsr_finish_sum(ep, sg.mode, sg.col0, i, j, ep.va ? ep.va[i] : 1.0, sg.vb ? sg.vb[j] : 1.0,
                       (ep.qa[i] * sg.qb[j]) * acc, mP, sg.mO, sg.ldm);
    }
}

}  // namespace

int nsr_launch_contract_finish(cudaStream_t st, const int32_t* tiles_dev, int64_t n_tiles, const ContractParams& ep,
                               const SegInfo& sg, int n_parts, const double* part, int64_t part_stride) {
    contract_finish_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(tiles_dev, ep, sg, n_parts, part, part_stride);
    NSR_CHECK(cudaGetLastError());
    return 0;
}

extern int nsr_umma_kblock;
// number of parts the kernel will really use for a request of n_parts over `cells` cells
int nsr_umma_parts(int64_t cells, int n_slices_a, int n_slices_b, int n_parts) {
    const int kb = (n_slices_a + n_slices_b > 6) ? 64 : nsr_umma_kblock;
    const int num_kb = (int)(cells / kb);
    if (n_parts <= 1 || num_kb < 1) return 1;
    const int per = (num_kb + n_parts - 1) / n_parts;
    return (num_kb + per - 1) / per;
}

int nsr_launch_contract_umma_splitk(nsr_ctx* ctx, cudaStream_t st, const int8_t* a, int64_t rows_a,
                                    int64_t rows_alloc_a, int n_slices_a, const NsrSegOperand* segs, int n_segs,
                                    int64_t n_pad, int n_slices_b, int wmax,
                                    const int32_t* tiles_dev, int64_t n_tiles, const ContractParams& ep,
                                    int64_t cell_begin, int64_t cell_end, const int* n_tiles_dev, int n_parts,
                                    double* part_out, int64_t part_stride);

int nsr_umma_stack = 1;      // test hook: stacked-B N = 256 MMAs in the single-CTA kernel
int nsr_epi_warps = 8;       // test hook: epilogue warps of the single-CTA kernel (8 or 16)
int nsr_umma_pair = 0;       // 0 -> single-CTA kernel (default: 3 % faster sustained), 1 -> cta_group::2 kernel
int nsr_epi_sleep_ns = 500;  // test hook: epilogue wait back-off
int nsr_umma_kblock = 128;   // test hook (nsr_set_option): 128 -> SWIZZLE_128B stages, 64 -> SWIZZLE_64B
int nsr_umma_dynamic = 1;    // 1: tiles claimed from a global counter, 0: static round-robin (single-CTA kernel)

#else   // NSR_SPLITK_TU
}  // namespace
extern int nsr_umma_kblock;
extern int nsr_epi_sleep_ns;
#endif

#if NSR_SPLITK_TU
int nsr_launch_contract_umma_splitk(nsr_ctx* ctx, cudaStream_t st, const int8_t* a, int64_t rows_a,
                                    int64_t rows_alloc_a, int n_slices_a, const NsrSegOperand* segs, int n_segs,
                                    int64_t n_pad, int n_slices_b, int wmax,
                                    const int32_t* tiles_dev, int64_t n_tiles, const ContractParams& ep,
                                    int64_t cell_begin, int64_t cell_end, const int* n_tiles_dev, int n_parts,
                                    double* part_out, int64_t part_stride) {
    if (n_tiles < 0 || n_parts < 2 || part_out == nullptr) {
        nsr_set_error("nsr_contract: the split-over-cells launch takes a plain tile list and at least two parts");
        return 2;
    }
#else
int nsr_launch_contract_umma(nsr_ctx* ctx, cudaStream_t st, const int8_t* a, int64_t rows_a,
                             int64_t rows_alloc_a, int n_slices_a, const NsrSegOperand* segs, int n_segs,
                             int64_t n_pad, int n_slices_b, int wmax,
                             const int32_t* tiles_dev, int64_t n_tiles, const ContractParams& ep,
                             int64_t cell_begin, int64_t cell_end, const int* n_tiles_dev, int n_parts,
                             double* part_out, int64_t part_stride) {
    if (n_parts > 1 && part_out != nullptr)
        return nsr_launch_contract_umma_splitk(ctx, st, a, rows_a, rows_alloc_a, n_slices_a, segs, n_segs, n_pad, n_slices_b,
                                               wmax, tiles_dev, n_tiles, ep, cell_begin, cell_end, n_tiles_dev, n_parts,
                                               part_out, part_stride);
    if (n_tiles < 0) {
        // pair-tile list (tile_row/2, tile_col, mask), -n_tiles entries: cta_group::2 kernel
        if (n_segs != 1 || n_slices_a != n_slices_b) {
            nsr_set_error("nsr_contract: the cta_group::2 engine takes one segment and equal plane counts");
            return 2;
        }
        const int n_slices = n_slices_a;
        CUtensorMap ma, mb;
        if (make_map(ctx, &ma, a, rows_a, rows_alloc_a, n_pad, n_slices, 128, NSR_TILE)) return 1;
        if (make_map(ctx, &mb, segs[0].slices, segs[0].info.rows_b, segs[0].rows_alloc, n_pad, n_slices, 128, NSR_TILE / 2)) return 1;
        UmmaArgs g;
        g.tiles = tiles_dev;
        g.n_tiles = (int)(-n_tiles);
        g.num_kb = (int)((cell_end - cell_begin) / 128);
        g.kb_begin = (int)(cell_begin / 128);
        g.stack_b = 0;
        g.tile_counter = nullptr;
        g.n_tiles_dev = nullptr;
        g.epi_sleep_ns = nsr_epi_sleep_ns;
        g.n_segs = 1;
        g.seg[0] = segs[0].info;
        g.ep = ep;
        if (n_slices == 3 && wmax == 4) return launch2<3, 4>(ctx, st, ma, mb, g);
        if (n_slices == 3 && wmax == 5) return launch2<3, 5>(ctx, st, ma, mb, g);
        if (n_slices == 4 && wmax == 5) return launch2<4, 5>(ctx, st, ma, mb, g);
        nsr_set_error("nsr_contract: unsupported (n_slices=%d, wmax=%d) for the tcgen05 pair engine", n_slices, wmax);
        return 2;
    }
#endif
    if (n_segs < 1 || n_segs > kMaxSegs) {
        nsr_set_error("nsr_contract: %d segments (1..%d)", n_segs, kMaxSegs);
        return 2;
    }
    // two stages of (SA + SB) planes x 128 rows x KB cells must fit the operand ring
    const int kb = (n_slices_a + n_slices_b > 6) ? 64 : nsr_umma_kblock;
    CUtensorMap ma;
    SegMaps mb;
    if (make_map(ctx, &ma, a, rows_a, rows_alloc_a, n_pad, n_slices_a, kb)) return 1;
    UmmaArgs g;
    for (int s = 0; s < n_segs; ++s) {
        if (make_map(ctx, &mb.b[s], segs[s].slices, segs[s].info.rows_b, segs[s].rows_alloc, n_pad, n_slices_b, kb)) return 1;
        g.seg[s] = segs[s].info;
    }
    for (int s = n_segs; s < kMaxSegs; ++s) { mb.b[s] = mb.b[0]; g.seg[s] = segs[0].info; }
    g.n_segs = n_segs;
    g.tiles = tiles_dev;
    g.n_tiles = (int)n_tiles;
    g.num_kb = (int)((cell_end - cell_begin) / kb);
    g.kb_begin = (int)(cell_begin / kb);
#if NSR_SPLITK_TU
    g.kb_per_part = (g.num_kb + n_parts - 1) / n_parts;
    g.n_parts = (g.num_kb + g.kb_per_part - 1) / g.kb_per_part;      // no empty part
    g.part_out = part_out;
    g.part_stride = part_stride;
#endif
    g.stack_b = (nsr_umma_stack != 0 && kb == 128) ? 1 : 0;
    g.tile_counter = nullptr;
    g.n_tiles_dev = n_tiles_dev;
    if (nsr_umma_dynamic && ctx->tile_counters) {
        g.tile_counter = ctx->tile_counters + (ctx->launch_seq++ % NSR_TILE_COUNTERS);
        NSR_CHECK(cudaMemsetAsync(g.tile_counter, 0, sizeof(int), st));
    }
    g.epi_sleep_ns = nsr_epi_sleep_ns;
    g.ep = ep;
    const int sa = n_slices_a, sb = n_slices_b;
    if (sa == 3 && sb == 3 && wmax == 4 && kb == 128) return launch<3, 3, 4, 128>(ctx, st, ma, mb, g);
    if (sa == 3 && sb == 3 && wmax == 5 && kb == 128) return launch<3, 3, 5, 128>(ctx, st, ma, mb, g);
    if (sa == 3 && sb == 3 && wmax == 4 && kb == 64) return launch<3, 3, 4, 64>(ctx, st, ma, mb, g);
    if (sa == 3 && sb == 3 && wmax == 5 && kb == 64) return launch<3, 3, 5, 64>(ctx, st, ma, mb, g);
    if (sa == 4 && sb == 4 && wmax == 5) return launch<4, 4, 5, 64>(ctx, st, ma, mb, g);
    if (sa == 1 && sb == 3 && wmax == 4 && kb == 128) return launch<1, 3, 4, 128>(ctx, st, ma, mb, g);
    if (sa == 1 && sb == 4 && wmax == 5 && kb == 128) return launch<1, 4, 5, 128>(ctx, st, ma, mb, g);
    if (sa == 1 && sb == 1 && wmax == 2 && kb == 128) return launch<1, 1, 2, 128>(ctx, st, ma, mb, g);
    nsr_set_error("nsr_contract: unsupported (n_slices_a=%d, n_slices_b=%d, wmax=%d, kblock=%d) for the tcgen05 engine",
                  sa, sb, wmax, kb);
    return 2;
}
