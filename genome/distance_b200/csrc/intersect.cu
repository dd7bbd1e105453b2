// intersect.cu -- kernel 4 (pairwise set intersection count) and kernel 5 (distance epilogue) of
// libgkd.so.  sm_100a only; no tensor cores (this is merge work, not a contraction).
//
// Reference semantics restated:
//   SequenceKmers.similarity(other): number of members of one HashSet<String> found in the other
//   SequenceKmers.distance(other):   I == 0 ? 1.0 : 1.0 - I / ((|A| + |B|) - I), |A|+|B| a Java int sum
//   (called at FastaDistanceProcessor.java:186, GenomeProcessor.java:140, DistanceRepsProcessor.java:101,190,
//    FastaDistanceRepsProcessor.java:128).
//   Pair order: strict upper triangle in list order (FastaDistanceProcessor.java:177) or every
//   (query, base) pair (GenomeProcessor.java:140-146).
//
// Kernel 4 design (B200), round 2: warp-cooperative merge over VALUE-partitioned runs.
//   Sets are sorted by a bijective mix of the key and carry a table of bucket offsets (gkd_internal.cuh),
//   so the merge partition is a table lookup: bucket f of A and bucket f of B cover the same key range.
//   * a warp owns a work item = (pair, run of 32-bucket groups); items come from one global counter and
//     are enumerated pair-major, so the warps resident at one time share a few row sets (served from the
//     126 MB L2) and walk a pair's two key streams front to back;
//   * per group, lane t reads its bucket bounds from both tables (two coalesced loads per set), lane 0
//     requests the group's two contiguous key ranges with TMA bulk copies (cp.async.bulk, SASS UBLKCP)
//     into the warp's own shared-memory stage guarded by an mbarrier (expect_tx / try_wait.parity); the
//     offsets of the group after next and the copies of the next group are in flight while the current
//     group is merged;
//   * every lane then merges its two runs (<= tmax keys each on average) from shared memory with a
//     branch-light two-pointer loop on the 32-bit low words (64-bit for keys wider than 42 bits): ~10
//     SASS instructions per step for 32 lanes, against ~100 per 58 keys for the round-1 ballot kernel;
//   * sets of different size meet at the level of the larger one: the lanes partition the larger set
//     exactly and several lanes share one (coarser) bucket of the smaller set, so a pair never costs
//     more than ~2x the keys of its larger set and no separate kernel is needed for skewed pairs;
//   * a group whose ranges do not fit the stage (rare: > 5 sigma, or mis-tuned tmax) is merged straight
//     from global memory by the same loop;
//   * counts: per-lane counters, one warp reduction (REDUX) and one atomicAdd per item.
// Bound: nominally HBM (algorithmic bytes 8 * (|A| + |B|) per pair); the stored bytes are
// ~4.25 per key (low word + table share), see DESIGN.md section 4 for the measured limiter.
#include <cstdlib>

#include "gkd_internal.cuh"

namespace gkd {

// ---- mbarrier / TMA / shared-memory wrappers --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (++spins > (1u << 24)) __trap();  // a lost copy must not hang the GPU
    } while (!done);
}

template <typename LowT>
__device__ __forceinline__ LowT lds_low(uint32_t addr);
template <>
__device__ __forceinline__ uint32_t lds_low<uint32_t>(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <>
__device__ __forceinline__ uint64_t lds_low<uint64_t>(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

// decode pair index t of this call into set ids
__device__ __forceinline__ void decode_pair(const PairSource &src, uint64_t t, uint32_t &ida, uint32_t &idb) {
    if (src.mode == PAIRS_UPPER) {
        upper_pair(src.first + t, src.n, ida, idb);
    } else if (src.mode == PAIRS_RECT) {
        ida = src.a[t / src.n];
        idb = src.b[t % src.n];
    } else if (src.mode == PAIRS_LIST_VS_ONE) {
        ida = src.a[t];
        idb = src.n;
    } else {
        ida = src.a[t];
        idb = src.b[t];
    }
}

// ---- the merge loops -----------------------------------------------------------------------------------
// Two-pointer count of equal keys of two ascending runs.  Equal heads advance both runs; the loop ends
// as soon as either run is exhausted (the rest of the other run cannot match).
// The shared-memory version is written in PTX to pin its shape: per step two compares, two predicated
// pointer bumps, a step counter, two predicated loads and the exit test (10 instructions for 32 lanes;
// the compiler's version of the same C loop took 13-15).  Every step advances one run, or both when the
// heads are equal, so matches = advances - steps.  The loads run ahead of the exit test and may read
// the word just past a run, which is still inside the CTA's shared memory.
template <typename LowT>
__device__ __forceinline__ uint32_t merge_smem(uint32_t pa, uint32_t ea, uint32_t pb, uint32_t eb);

#ifndef GKD_MERGE_UNROLL2
#define GKD_MERGE_UNROLL2 1
#endif
#ifndef GKD_MERGE_ONE_LDS
#define GKD_MERGE_ONE_LDS 0  // measured on B200 (300 genomes): 345 ms against 284 ms for the two-load step
#endif

#if GKD_MERGE_ONE_LDS
// Experiment kept for the record: one shared-memory load per step (exactly one run advances; ties advance A and
// are counted there), a single LDS with a selected address serving every lane.  It cuts the shared-memory
// wavefronts (the kernel's co-limiter) by ~40 % but needs 10.5 instead of 8 instructions per step and puts two
// selects on the load's dependency chain: slower overall (issue-bound), so the two-load step below is the default.
#define GKD_MERGE_STEP(T, SZ)               \
    "setp.le." T " ple, a, b;\n"           \
    "setp.eq." T " peq, a, b;\n"           \
    "@peq add.u32 %2, %2, 1;\n"            \
    "@ple add.u32 %0, %0, " SZ ";\n"       \
    "@!ple add.u32 %1, %1, " SZ ";\n"      \
    "selp.u32 adr, %0, %1, ple;\n"         \
    "ld.shared." T " v, [adr];\n"          \
    "selp." T " a, v, a, ple;\n"           \
    "selp." T " b, b, v, ple;\n"
#define GKD_MERGE_REGS(T) ".reg .pred ple, peq, pgo;\n.reg ." T " a, b, v;\n.reg .u32 adr;\n"
#define GKD_MERGE_COUNT2 ""
#define GKD_MERGE_COUNT1 ""
#else
#define GKD_MERGE_STEP(T, SZ)               \
    "setp.gt." T " pgt, a, b;\n"           \
    "setp.lt." T " plt, a, b;\n"           \
    "@!pgt add.u32 %0, %0, " SZ ";\n"      \
    "@!plt add.u32 %1, %1, " SZ ";\n"      \
    "@!pgt ld.shared." T " a, [%0];\n"     \
    "@!plt ld.shared." T " b, [%1];\n"
#define GKD_MERGE_REGS(T) ".reg .pred pgt, plt, pgo;\n.reg ." T " a, b;\n"
#define GKD_MERGE_COUNT2 "add.u32 %2, %2, 2;\n"
#define GKD_MERGE_COUNT1 "add.u32 %2, %2, 1;\n"
#endif

// With GKD_MERGE_UNROLL2 the loop runs two steps per trip while both runs hold at least two more keys (no exit
// test between them), then finishes with the one-step loop.
#define GKD_MERGE_BODY(T, SZ, TAG)                                                          \
    "{\n"                                                                                   \
    GKD_MERGE_REGS(T)                                                                       \
    "ld.shared." T " a, [%0];\n"                                                            \
    "ld.shared." T " b, [%1];\n"                                                            \
    "setp.lt.u32 pgo, %0, %5;\n"                                                            \
    "setp.lt.and.u32 pgo, %1, %6, pgo;\n"                                                   \
    "@!pgo bra " TAG "_ONE_%=;\n"                                                           \
    TAG "_TWO_%=:\n"                                                                        \
    GKD_MERGE_STEP(T, SZ)                                                                   \
    GKD_MERGE_STEP(T, SZ)                                                                   \
    GKD_MERGE_COUNT2                                                                        \
    "setp.lt.u32 pgo, %0, %5;\n"                                                            \
    "setp.lt.and.u32 pgo, %1, %6, pgo;\n"                                                   \
    "@pgo bra " TAG "_TWO_%=;\n"                                                            \
    "setp.lt.u32 pgo, %0, %3;\n"                                                            \
    "setp.lt.and.u32 pgo, %1, %4, pgo;\n"                                                   \
    "@!pgo bra " TAG "_END_%=;\n"                                                           \
    TAG "_ONE_%=:\n"                                                                        \
    GKD_MERGE_STEP(T, SZ)                                                                   \
    GKD_MERGE_COUNT1                                                                        \
    "setp.lt.u32 pgo, %0, %3;\n"                                                            \
    "setp.lt.and.u32 pgo, %1, %4, pgo;\n"                                                   \
    "@pgo bra " TAG "_ONE_%=;\n"                                                            \
    TAG "_END_%=:\n"                                                                        \
    "}\n"

// what merge_smem returns from its three in/out operands (pa, pb, and the step counter or the match counter)
__device__ __forceinline__ uint32_t merge_result(uint32_t pa, uint32_t pb, uint32_t start, uint32_t counter, int shift) {
#if GKD_MERGE_ONE_LDS
    (void)pa, (void)pb, (void)start, (void)shift;
    return counter;  // matches counted directly
#else
    return ((pa + pb - start) >> shift) - counter;  // every step advances one run, or both on a match
#endif
}

template <>
__device__ __forceinline__ uint32_t merge_smem<uint32_t>(uint32_t pa, uint32_t ea, uint32_t pb, uint32_t eb) {
    if (pa >= ea || pb >= eb) return 0;
    const uint32_t start = pa + pb;
    uint32_t counter = 0;
    // two unchecked steps need two keys left in each run: pa < ea - 4 and pb < eb - 4 (the runs are not empty here)
    const uint32_t ea2 = GKD_MERGE_UNROLL2 ? ea - 4u : 0u, eb2 = GKD_MERGE_UNROLL2 ? eb - 4u : 0u;
    asm volatile(GKD_MERGE_BODY("u32", "4", "M32") : "+r"(pa), "+r"(pb), "+r"(counter) : "r"(ea), "r"(eb), "r"(ea2), "r"(eb2) : "memory");
    return merge_result(pa, pb, start, counter, 2);
}

template <>
__device__ __forceinline__ uint32_t merge_smem<uint64_t>(uint32_t pa, uint32_t ea, uint32_t pb, uint32_t eb) {
    if (pa >= ea || pb >= eb) return 0;
    const uint32_t start = pa + pb;
    uint32_t counter = 0;
    const uint32_t ea2 = GKD_MERGE_UNROLL2 ? ea - 8u : 0u, eb2 = GKD_MERGE_UNROLL2 ? eb - 8u : 0u;
    asm volatile(GKD_MERGE_BODY("u64", "8", "M64") : "+r"(pa), "+r"(pb), "+r"(counter) : "r"(ea), "r"(eb), "r"(ea2), "r"(eb2) : "memory");
    return merge_result(pa, pb, start, counter, 3);
}

template <typename LowT>
__device__ __forceinline__ uint32_t merge_global(const LowT *__restrict__ pa, const LowT *__restrict__ ea,
                                                 const LowT *__restrict__ pb, const LowT *__restrict__ eb) {
    uint32_t cnt = 0;
    if (pa < ea && pb < eb) {
        LowT a = __ldg(pa), b = __ldg(pb);
        for (;;) {
            const bool le = a <= b, ge = b <= a;
            cnt += (le && ge) ? 1u : 0u;
            if (le) pa++;
            if (ge) pb++;
            if (pa >= ea || pb >= eb) break;
            if (le) a = __ldg(pa);
            if (ge) b = __ldg(pb);
        }
    }
    return cnt;
}

// bounds [lo, hi) of lane bucket f (at walk level L) in a set whose table is at S.level
__device__ __forceinline__ void lane_run(const SubSet &S, uint32_t L, uint32_t f, bool active, uint32_t &lo,
                                         uint32_t &hi) {
    lo = hi = 0;
    if (!active) return;
    if (S.level >= L) {
        const uint32_t sh = S.level - L;
        lo = __ldg(S.offs + ((size_t)f << sh));
        hi = __ldg(S.offs + ((size_t)(f + 1) << sh));
    } else {  // coarser table: the lanes of 2^sh consecutive buckets share one bucket of this set
        const uint32_t c = f >> (L - S.level);
        lo = __ldg(S.offs + c);
        hi = __ldg(S.offs + c + 1);
    }
}

// Kernel configuration: CAP bytes per set per stage, STAGES stages per warp, WARPS warps per CTA,
// CTAS CTAs per SM.
template <int CAP_, int STAGES_, int WARPS_, int CTAS_>
struct BucketCfg {
    static constexpr int CAP = CAP_, STAGES = STAGES_, WARPS = WARPS_, CTAS = CTAS_;
    static constexpr uint32_t STAGE_BYTES = 2u * CAP;                      // A range + B range
    static constexpr uint32_t WARP_BYTES = STAGES * STAGE_BYTES;
    static constexpr uint32_t BAR_OFF = WARPS * WARP_BYTES;
    static constexpr uint32_t SMEM = BAR_OFF + WARPS * STAGES * 8u;
    static_assert(CAP % 16 == 0, "stages hold whole 16-byte TMA units");
    static_assert((size_t)(SMEM + 1024) * CTAS <= 227u * 1024u, "does not fit the SM");
    static_assert(WARPS * CTAS <= 64, "at most 64 warps per SM");
};

struct OffsRegs {  // bucket bounds of one lane for one group (key indices)
    uint32_t loA, hiA, loB, hiB;
};
struct StageRegs {  // where this lane's runs of one group are, once staged
    uint32_t pa, ea, pb, eb;  // shared-memory addresses (staged) or key indices (direct)
    uint32_t mode;            // 0 = nothing to merge, 1 = staged and a copy is in flight, 2 = direct from global
};

template <typename LowT, class C>
__global__ void __launch_bounds__(C::WARPS * 32, C::CTAS)
    k_intersect_bucket(const SetDesc *__restrict__ sets, PairSource src, int use_pal, IsectPlan plan,
                       uint32_t *__restrict__ counts, unsigned long long *__restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t KEYS16 = 16u / (uint32_t)sizeof(LowT);  // keys per 16-byte TMA unit
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t buf0 = smem_u32(smem_raw) + warp * C::WARP_BYTES;
    const uint32_t bar0 = smem_u32(smem_raw) + C::BAR_OFF + warp * (C::STAGES * 8u);
    if (lane == 0) {
        for (int i = 0; i < C::STAGES; i++) mbar_init(bar0 + i * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phases = 0;  // bit s = parity to wait for on the barrier of stage s

    const uint64_t total_items = (src.count_ptr ? (uint64_t)*src.count_ptr : src.count) * (uint64_t)plan.items_per_pair;
    unsigned long long next_item = 0;
    if (lane == 0) next_item = atomicAdd(work_counter, 1ull);
    for (;;) {
        const uint64_t item = __shfl_sync(0xffffffffu, next_item, 0);
        if (item >= total_items) break;
        if (lane == 0) next_item = atomicAdd(work_counter, 1ull);  // fetched while this item is processed
        const uint64_t pair = item / plan.items_per_pair;
        const uint32_t chunk = (uint32_t)(item - pair * plan.items_per_pair);
        uint32_t ida, idb;
        decode_pair(src, pair, ida, idb);
        const SetDesc *da = sets + ida, *db = sets + idb;
        const SubSet SA = use_pal ? da->pal : da->main, SB = use_pal ? db->pal : db->main;
        if (SA.n == 0 || SB.n == 0) continue;
        // walk level: the larger set averages <= tmax keys per lane bucket
        const uint32_t nmax = SA.n > SB.n ? SA.n : SB.n, lmax = SA.level > SB.level ? SA.level : SB.level;
        uint32_t L = level_for(nmax, plan.tmax);
        if (L < plan.level_min) L = plan.level_min;
        if (L > lmax) L = lmax;
        const uint32_t n_buckets = 1u << L;
        const uint32_t n_groups = (n_buckets + 31u) >> 5;
        const uint32_t g0 = chunk * plan.groups_per_item;
        if (g0 >= n_groups) continue;
        const uint32_t g1 = (n_groups - g0 > plan.groups_per_item) ? g0 + plan.groups_per_item : n_groups;
        const LowT *lowsA = (const LowT *)SA.lows, *lowsB = (const LowT *)SB.lows;

        // phase O: this lane's bucket bounds of group g
        auto load_offs = [&](uint32_t g) {
            OffsRegs o;
            const uint32_t f = g * 32u + lane;
            const bool active = f < n_buckets;
            lane_run(SA, L, f, active, o.loA, o.hiA);
            lane_run(SB, L, f, active, o.loB, o.hiB);
            return o;
        };
        // phase T: request the group's two key ranges into stage s (or decide to merge from global)
        auto stage_group = [&](const OffsRegs &o, uint32_t g, uint32_t s) {
            StageRegs r;
            const uint32_t last = (n_buckets - g * 32u >= 32u) ? 31u : (n_buckets - g * 32u - 1u);
            const uint32_t a_lo = __shfl_sync(0xffffffffu, o.loA, 0), a_hi = __shfl_sync(0xffffffffu, o.hiA, last);
            const uint32_t b_lo = __shfl_sync(0xffffffffu, o.loB, 0), b_hi = __shfl_sync(0xffffffffu, o.hiB, last);
            const uint32_t a0 = a_lo & ~(KEYS16 - 1u), b0 = b_lo & ~(KEYS16 - 1u);
            const uint32_t bytesA = a_hi > a_lo ? (uint32_t)align16((uint64_t)(a_hi - a0) * sizeof(LowT)) : 0u;
            const uint32_t bytesB = b_hi > b_lo ? (uint32_t)align16((uint64_t)(b_hi - b0) * sizeof(LowT)) : 0u;
            if (bytesA == 0 || bytesB == 0) {  // one side has no key in this group: nothing can match
                r.pa = r.ea = r.pb = r.eb = 0;
                r.mode = 0;
            } else if (bytesA > (uint32_t)C::CAP || bytesB > (uint32_t)C::CAP) {
                r.pa = o.loA, r.ea = o.hiA, r.pb = o.loB, r.eb = o.hiB;
                r.mode = 2;
            } else {
                const uint32_t bufA = buf0 + s * C::STAGE_BYTES, bufB = bufA + C::CAP, bar = bar0 + s * 8u;
                if (lane == 0) {
                    mbar_expect_tx(bar, bytesA + bytesB);
                    tma_load_1d(bufA, lowsA + a0, bytesA, bar);
                    tma_load_1d(bufB, lowsB + b0, bytesB, bar);
                }
                r.pa = bufA + (o.loA - a0) * (uint32_t)sizeof(LowT);
                r.ea = bufA + (o.hiA - a0) * (uint32_t)sizeof(LowT);
                r.pb = bufB + (o.loB - b0) * (uint32_t)sizeof(LowT);
                r.eb = bufB + (o.hiB - b0) * (uint32_t)sizeof(LowT);
                r.mode = 1;
            }
            return r;
        };
        // phase M: wait for the stage and merge this lane's runs
        auto merge_group = [&](const StageRegs &r, uint32_t s) -> uint32_t {
            uint32_t c = 0;
            if (r.mode == 1) {
                mbar_wait(bar0 + s * 8u, (phases >> s) & 1u);
                phases ^= 1u << s;
                c = merge_smem<LowT>(r.pa, r.ea, r.pb, r.eb);
            } else if (r.mode == 2) {
                c = merge_global<LowT>(lowsA + r.pa, lowsA + r.ea, lowsB + r.pb, lowsB + r.eb);
            }
            __syncwarp();  // every lane is done reading the stage before it is requested again
            return c;
        };

        uint32_t cnt = 0, s = 0;
        OffsRegs on = load_offs(g0);
        StageRegs cur = stage_group(on, g0, 0);
        if (g0 + 1 < g1) on = load_offs(g0 + 1);
        for (uint32_t g = g0; g < g1; g++) {
            if (C::STAGES >= 2) {
                StageRegs nxt = cur;
                if (g + 1 < g1) {
                    nxt = stage_group(on, g + 1, s ^ 1u);
                    if (g + 2 < g1) on = load_offs(g + 2);
                }
                cnt += merge_group(cur, s);
                cur = nxt;
                s ^= 1u;
            } else {
                cnt += merge_group(cur, 0);
                if (g + 1 < g1) {
                    cur = stage_group(on, g + 1, 0);
                    if (g + 2 < g1) on = load_offs(g + 2);
                }
            }
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&counts[pair], cnt);
    }
}

// Configurations (index selected with GKD_ISECT_CFG; see profiles/ for the sweep).  The stage must hold
// 32 lanes x tmax keys plus ~5 sigma of the bucket-occupancy noise and the 16-byte alignment slack.
using BCfg0 = BucketCfg<2560, 2, 4, 5>;   // tmax 16 (u32) / 8 (u64), double-buffered, 20 warps per SM
using BCfg1 = BucketCfg<2560, 1, 4, 10>;  // single stage, 40 warps per SM
using BCfg2 = BucketCfg<3840, 2, 4, 3>;   // tmax 24 / 12, double-buffered, 12 warps per SM
using BCfg3 = BucketCfg<3840, 1, 4, 7>;   // single stage, 28 warps per SM
using BCfg4 = BucketCfg<1536, 2, 4, 9>;   // tmax 8 / 4, double-buffered, 36 warps per SM
using BCfg5 = BucketCfg<1536, 1, 4, 16>;  // single stage, 64 warps per SM
using BCfg6 = BucketCfg<3072, 1, 4, 9>;   // tmax 24 with a tight stage, 36 warps per SM
using BCfg7 = BucketCfg<3328, 1, 4, 8>;   // tmax 24, 32 warps per SM
using BCfg8 = BucketCfg<5632, 1, 4, 4>;   // tmax 48 / 24, 16 warps per SM
using BCfg9 = BucketCfg<5632, 1, 2, 9>;   // tmax 48 / 24, 18 warps per SM
constexpr int N_BCFG = 10;
constexpr int DEFAULT_BCFG = 9;
static const uint32_t g_cfg_tmax32[N_BCFG] = {16, 16, 24, 24, 8, 8, 24, 24, 48, 48};

#define GKD_FOR_EACH_BCFG(X) \
    X(0, BCfg0) X(1, BCfg1) X(2, BCfg2) X(3, BCfg3) X(4, BCfg4) X(5, BCfg5) X(6, BCfg6) X(7, BCfg7) X(8, BCfg8) X(9, BCfg9)

static int env_cfg() {
    const char *e = getenv("GKD_ISECT_CFG");
    int c = e ? atoi(e) : DEFAULT_BCFG;
    return (c < 0 || c >= N_BCFG) ? DEFAULT_BCFG : c;
}

cudaError_t intersect_configure() {
    cudaError_t e;
#define X(i, C)                                                                                                          \
    if ((e = cudaFuncSetAttribute(k_intersect_bucket<uint32_t, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM))) \
        return e;                                                                                                        \
    if ((e = cudaFuncSetAttribute(k_intersect_bucket<uint64_t, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM))) \
        return e;
    GKD_FOR_EACH_BCFG(X)
#undef X
    return cudaSuccess;
}

uint32_t intersect_default_tmax(int low_bits) {
    const char *e = getenv("GKD_ISECT_TMAX");
    uint32_t t = e ? (uint32_t)atoi(e) : g_cfg_tmax32[env_cfg()];
    if (low_bits == 64 && !e) t /= 2;
    return t < 1 ? 1 : t;
}

uint32_t intersect_warps_per_sm(int) {
    const int c = env_cfg();
#define X(i, C) \
    if (c == i) return C::WARPS * C::CTAS;
    GKD_FOR_EACH_BCFG(X)
#undef X
    return 16;
}

template <typename LowT, class C>
static cudaError_t launch_bcfg(const SetDesc *sets, PairSource src, int use_pal, const IsectPlan &plan, uint32_t *counts,
                               unsigned long long *work_counter, int n_sms, cudaStream_t s) {
    const uint64_t items = src.count * (uint64_t)plan.items_per_pair;
    uint64_t grid = (uint64_t)n_sms * C::CTAS;  // persistent: every SM holds CTAS resident CTAs
    const uint64_t need = (items + C::WARPS - 1) / C::WARPS;
    if (grid > need) grid = need;
    k_intersect_bucket<LowT, C><<<(unsigned)grid, C::WARPS * 32, C::SMEM, s>>>(sets, src, use_pal, plan, counts, work_counter);
    return cudaGetLastError();
}

cudaError_t launch_intersect(const SetDesc *sets, PairSource src, int use_pal, const IsectPlan &plan, uint32_t *counts,
                             unsigned long long *work_counter, int n_sms, cudaStream_t s) {
    if (src.count == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    const int c = env_cfg();
#define X(i, C)                                                                                               \
    if (c == i)                                                                                               \
        return plan.low_bits == 32 ? launch_bcfg<uint32_t, C>(sets, src, use_pal, plan, counts, work_counter, n_sms, s) \
                                   : launch_bcfg<uint64_t, C>(sets, src, use_pal, plan, counts, work_counter, n_sms, s);
    GKD_FOR_EACH_BCFG(X)
#undef X
    return cudaErrorInvalidValue;
}

// ---- kernel 5: distance epilogue ----------------------------------------------------------------------
// SequenceKmers.similarity / distance of one pair from the canonical counts (c = |C_A n C_B|, cp = palindromic part)
__device__ __forceinline__ void pair_result(const SetDesc *__restrict__ sets, uint32_t ida, uint32_t idb, uint64_t c, uint64_t cp,
                                            int both_strands, uint64_t &I, uint64_t &sa, uint64_t &sb, double &dist) {
    const uint32_t nA = sets[ida].main.n, nB = sets[idb].main.n;
    const uint32_t pA = sets[ida].pal.n, pB = sets[idb].pal.n;
    if (both_strands) {
        // the reference's sets hold both strands: |S| = 2|C| - P, I = 2|C_A n C_B| - P(C_A n C_B)
        I = 2 * c - cp;
        sa = 2ull * nA - pA;
        sb = 2ull * nB - pB;
    } else {
        I = c;
        sa = nA;
        sb = nB;
    }
    dist = 1.0;
    const double similarity = (double)I;
    if (similarity > 0) {
        // (this.size() + other.size()) is a Java int addition
        const int32_t sum = (int32_t)((uint32_t)sa + (uint32_t)sb);
        const double uni = (double)sum - similarity;
        dist = 1.0 - similarity / uni;
    }
}

__global__ void __launch_bounds__(256)
    k_epilogue(const SetDesc *__restrict__ sets, PairSource src, const uint32_t *__restrict__ counts,
               const uint32_t *__restrict__ pal_counts, int both_strands, EpilogueOut out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= src.count) return;
    uint32_t ida, idb;
    decode_pair(src, t, ida, idb);
    uint64_t I, sa, sb;
    double d;
    pair_result(sets, ida, idb, counts[t], pal_counts ? pal_counts[t] : 0, both_strands, I, sa, sb, d);
    if (out.inter) out.inter[t] = I;
    if (out.dist) out.dist[t] = d;
    if (out.contain_a) out.contain_a[t] = sa ? (double)I / (double)sa : 0.0;
    if (out.contain_b) out.contain_b[t] = sb ? (double)I / (double)sb : 0.0;
}

// Greedy representative pass (DistanceRepsProcessor.java:185-201, FastaDistanceRepsProcessor.java:124-146): the
// candidate was intersected with every current representative (counts[t] for reps[t]); it joins the list unless
// one of them is within max_dist.  The list and its length live on the device, so the host queues the launches
// of every candidate back to back without reading anything.
__global__ void __launch_bounds__(256)
    k_greedy_decide(const SetDesc *__restrict__ sets, uint32_t *__restrict__ reps, uint32_t *__restrict__ n_reps, uint32_t cand,
                    uint32_t visit, uint32_t *__restrict__ counts, uint32_t *__restrict__ pal_counts, int both_strands,
                    double max_dist, uint8_t *__restrict__ is_rep) {
    __shared__ int s_found;
    if (threadIdx.x == 0) s_found = 0;
    __syncthreads();
    const uint32_t n = *n_reps;
    int found = 0;
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
        uint64_t I, sa, sb;
        double d;
        pair_result(sets, reps[t], cand, counts[t], pal_counts ? pal_counts[t] : 0, both_strands, I, sa, sb, d);
        found |= d <= max_dist ? 1 : 0;
        counts[t] = 0;  // ready for the next candidate
        if (pal_counts) pal_counts[t] = 0;
    }
    if (found) atomicOr(&s_found, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        is_rep[visit] = s_found ? 0 : 1;
        if (!s_found) {
            reps[n] = cand;
            *n_reps = n + 1;
        }
    }
}

cudaError_t launch_greedy_decide(const SetDesc *sets, uint32_t *reps, uint32_t *n_reps, uint32_t cand, uint32_t visit,
                                 uint32_t *counts, uint32_t *pal_counts, int both_strands, double max_dist, uint8_t *is_rep,
                                 cudaStream_t s) {
    k_greedy_decide<<<1, 256, 0, s>>>(sets, reps, n_reps, cand, visit, counts, pal_counts, both_strands, max_dist, is_rep);
    return cudaGetLastError();
}

cudaError_t launch_epilogue(const SetDesc *sets, PairSource src, const uint32_t *counts, const uint32_t *pal_counts,
                            int both_strands, EpilogueOut out, cudaStream_t s) {
    if (src.count == 0) return cudaSuccess;
    uint64_t blocks = (src.count + 255) / 256;
    k_epilogue<<<(unsigned)blocks, 256, 0, s>>>(sets, src, counts, pal_counts, both_strands, out);
    return cudaGetLastError();
}

}  // namespace gkd
