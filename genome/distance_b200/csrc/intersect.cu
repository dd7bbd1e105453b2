// intersect.cu -- kernel 4 (pairwise sorted-set intersection count) and kernel 5 (distance epilogue)
// of libgkd.so.  sm_100a only; no tensor cores (this is merge work, not a contraction).
//
// Reference semantics restated:
//   SequenceKmers.similarity(other): number of members of one HashSet<String> found in the other
//   SequenceKmers.distance(other):   I == 0 ? 1.0 : 1.0 - I / ((|A| + |B|) - I), |A|+|B| a Java int sum
//   (called at FastaDistanceProcessor.java:186, GenomeProcessor.java:140, DistanceRepsProcessor.java:101,190,
//    FastaDistanceRepsProcessor.java:128).
//   Pair order: strict upper triangle in list order (FastaDistanceProcessor.java:177) or every
//   (query, base) pair (GenomeProcessor.java:140-146).
//
// Kernel 4 design (B200):
//   * persistent CTAs pull work items (pair, merge-path segment) from one global counter; pairs are
//     enumerated row-major so CTAs resident at the same time mostly share the row genome, which is
//     then served from the 126 MB L2 instead of HBM (measured DRAM traffic ~0.47x algorithmic);
//   * each input is streamed through a shared-memory ring filled by TMA bulk copies
//     (cp.async.bulk, 4 KiB blocks, one mbarrier per ring slot, thread 0 issues);
//     sets carry a sentinel tail so the merge needs no bounds checks; a slot is refilled as soon
//     as its block is consumed, which keeps 0.3-0.75 windows of loads in flight per CTA;
//   * a round merges W = THREADS x VT keys: every thread finds its merge-path split in shared
//     memory (byte-address binary search) and then merges VT keys serially, counting equal heads;
//     the last thread's end point advances the stream heads;
//   * segment starts inside a pair are found by a warp-cooperative 32-ary merge-path search on
//     global memory (__ballot_sync / __popc select the sub-range);
//   * per-warp shuffle reduction, one atomicAdd per warp per item.
// Bound: nominally HBM; measured limiter is the shared-memory pipe + issue slots (see profiles/).
// Algorithmic bytes: 8 * (|A| + |B|) per pair.
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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
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
    } else {
        ida = src.a[t];
        idb = src.b[t];
    }
}

// Merge-path split of diagonal d over two global arrays (A first on ties): the number of A keys
// among the first d merged keys.  Warp-cooperative 32-ary search; all lanes return the result.
__device__ __forceinline__ uint32_t diag_search_global(const uint64_t *__restrict__ A, uint32_t nA,
                                                       const uint64_t *__restrict__ B, uint32_t nB, uint64_t d) {
    const int lane = threadIdx.x & 31;
    uint32_t lo = d > nB ? (uint32_t)(d - nB) : 0u;
    uint32_t hi = d < nA ? (uint32_t)d : nA;
    while (lo < hi) {
        uint32_t span = hi - lo;
        uint32_t step = (span + 31) / 32;
        uint32_t c = lo + lane * step;
        bool pred = false;
        if (c < hi) pred = A[c] <= B[d - 1 - c];
        uint32_t trues = __popc(__ballot_sync(0xffffffffu, pred));  // pred is monotone: a prefix of lanes
        uint32_t nlo = trues ? lo + (trues - 1) * step + 1 : lo;
        uint32_t nhi = lo + trues * step;
        if (nhi > hi) nhi = hi;
        if (trues == 0) nhi = lo;
        lo = nlo;
        hi = nhi < nlo ? nlo : nhi;
    }
    return lo;
}

constexpr int WK_WARPS = 8;  // warp-cooperative variant: warps per CTA
#ifndef GKD_WK_CTAS
#define GKD_WK_CTAS 3
#endif
#ifndef GKD_WK_BLK
#define GKD_WK_BLK 256
#endif
constexpr int WK_CTAS = GKD_WK_CTAS;  // CTAs per SM

// Kernel configuration: THREADS x VT keys per round, ring of NBLK blocks per input.
template <int THREADS_, int VT_, int NBLK_, int CTAS_>
struct IsectCfg {
    static constexpr int THREADS = THREADS_, VT = VT_, NBLK = NBLK_, CTAS = CTAS_;
    static constexpr int W = THREADS * VT;
    static constexpr int CAP = NBLK * ISECT_BLK;         // keys per ring
    static constexpr uint32_t CAPB = (uint32_t)CAP * 8u;  // bytes per ring
    static constexpr uint32_t BAR_OFF = 2u * CAPB;
    static constexpr uint32_t MISC_OFF = BAR_OFF + 2u * NBLK * 8u;
    static constexpr uint32_t SMEM = MISC_OFF + 64u;
    static_assert(W <= ISECT_W_MAX, "sets are padded for windows up to ISECT_W_MAX keys");
    static constexpr int WIN = W;  // furthest key index a round may read, relative to the head
    static_assert(CAP >= WIN + 1 + ISECT_BLK, "ring must hold a full round window at any alignment");
    static_assert((size_t)SMEM * CTAS <= 227u * 1024u - 1024u * CTAS, "does not fit the SM");
};

struct IsectMisc {  // broadcast slots (double-buffered by item / round parity)
    unsigned long long item[2];
    uint32_t ida[2], idb[2];
    uint32_t heads[2];
    uint32_t start[2];
};

// One input stream: a ring of NBLK blocks of a sentinel-padded sorted key array.  Every thread keeps
// an identical copy of this state; only thread 0 issues copies.
struct Stream {
    const uint64_t *keys;
    uint32_t head;    // next unconsumed key (set position)
    uint32_t hidx;    // ring index (keys) of `head`
    uint32_t g0, v0;  // first block of this item and its CTA-lifetime virtual block number
    uint32_t issued;  // next block (set numbering) to request
    uint32_t ready;   // blocks < ready are known to have landed
    uint32_t rbar;    // shared address of the mbarrier of block `ready`
    uint32_t rpar;    // phase parity to wait for on that barrier
    uint32_t limit;   // one past the last block this item can touch
};

template <class C>
__device__ __forceinline__ void stream_begin(Stream &s, const uint64_t *keys, uint32_t n, uint32_t head0,
                                             uint32_t max_consume, uint32_t vnext) {
    s.keys = keys;
    s.head = head0;
    s.g0 = head0 / ISECT_BLK;
    s.v0 = vnext;
    s.hidx = (vnext % C::NBLK) * ISECT_BLK + head0 % ISECT_BLK;
    s.issued = s.g0;
    s.ready = s.g0;
    s.rbar = (vnext % C::NBLK) * 8u;  // offset; the array base is added at the wait
    s.rpar = (vnext / C::NBLK) & 1u;
    uint64_t last_pos = (uint64_t)head0 + max_consume;
    if (last_pos > n) last_pos = n;
    s.limit = (uint32_t)((last_pos + C::WIN) / ISECT_BLK) + 1;
}

// request every block whose ring slot is free
template <class C>
__device__ __forceinline__ void stream_issue(Stream &s, uint32_t ring_addr, uint32_t bar_addr) {
    uint32_t upto = s.head / ISECT_BLK + C::NBLK;
    if (upto > s.limit) upto = s.limit;
    if (threadIdx.x == 0) {
        for (uint32_t g = s.issued; g < upto; g++) {
            uint32_t slot = (s.v0 + (g - s.g0)) % C::NBLK;
            uint32_t bar = bar_addr + slot * 8;
            mbar_expect_tx(bar, ISECT_BLK * 8);
            tma_load_1d(ring_addr + slot * (ISECT_BLK * 8), s.keys + (size_t)g * ISECT_BLK, ISECT_BLK * 8, bar);
        }
    }
    if (upto > s.issued) s.issued = upto;
}

// every thread blocks until blocks [ready, upto) have landed (observing the mbarrier phase is what
// makes the async-proxy writes visible to that thread)
template <class C>
__device__ __forceinline__ void stream_wait(Stream &s, uint32_t upto, uint32_t bar_addr) {
    while (s.ready < upto) {
        mbar_wait(bar_addr + s.rbar, s.rpar);
        s.ready++;
        s.rbar += 8u;
        if (s.rbar == C::NBLK * 8u) {
            s.rbar = 0;
            s.rpar ^= 1u;
        }
    }
}

// One merge round over r <= W keys.  FULL rounds (r == W) run the branch-free VT-step merge.
template <class C, bool FULL>
__device__ __forceinline__ uint32_t merge_round(uint32_t hA, uint32_t hB, uint32_t rA0, uint32_t rB0, uint32_t r,
                                                uint32_t &cnt, bool &is_last) {
    constexpr uint32_t CAPB = C::CAPB;
    const uint32_t rA1 = rA0 + CAPB, rB1 = rB0 + CAPB;
    uint32_t d = (uint32_t)threadIdx.x * C::VT;
    if (!FULL && d > r) d = r;
    // merge-path split of this thread's diagonal inside the window (A first on ties)
    uint32_t lo = 0, hi = d;
    const uint32_t bBase = hB + (d - 1) * 8u;  // address of B[d-1] before wrapping (unused when d == 0)
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        uint32_t aa = hA + mid * 8u;
        if (aa >= rA1) aa -= CAPB;
        uint32_t bb = bBase - mid * 8u;
        if (bb >= rB1) bb -= CAPB;
        if (lds64(aa) <= lds64(bb)) lo = mid + 1;
        else hi = mid;
    }
    uint32_t pa = hA + lo * 8u;
    if (pa >= rA1) pa -= CAPB;
    uint32_t pb = hB + (d - lo) * 8u;
    if (pb >= rB1) pb -= CAPB;
    const uint32_t pa0 = pa;
    uint64_t a = lds64(pa), b = lds64(pb);
    uint32_t steps = C::VT;
    if (!FULL) {
        steps = r - d;
        if (steps > (uint32_t)C::VT) steps = C::VT;
    }
#pragma unroll
    for (int s = 0; s < C::VT; s++) {
        if (FULL || (uint32_t)s < steps) {
            const bool take_a = a <= b;
            cnt += (a == b) ? 1u : 0u;  // counted once, when the A copy is consumed
            if (take_a) {
                pa += 8u;
                if (pa == rA1) pa = rA0;
                a = lds64(pa);
            } else {
                pb += 8u;
                if (pb == rB1) pb = rB0;
                b = lds64(pb);
            }
        }
    }
    // the thread that ends exactly on the round's last diagonal reports how many A keys were used
    is_last = FULL ? (threadIdx.x == C::THREADS - 1) : (d < r && d + C::VT >= r);
    uint32_t usedA = (pa >= pa0 ? pa - pa0 : pa + CAPB - pa0) / 8u;
    return lo + usedA;
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
    k_intersect(const SetDesc *__restrict__ sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                uint32_t *__restrict__ counts, unsigned long long *__restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem0 = smem_u32(smem_raw);
    const uint32_t ringA = smem0, ringB = smem0 + C::CAPB;
    const uint32_t barA = smem0 + C::BAR_OFF, barB = barA + C::NBLK * 8u;
    IsectMisc &misc = *reinterpret_cast<IsectMisc *>(smem_raw + C::MISC_OFF);
    const int tid = threadIdx.x;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < C::NBLK; s++) {
            mbar_init(barA + s * 8, 1);
            mbar_init(barB + s * 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint64_t total_items = src.count * (uint64_t)max_segs;
    uint32_t vnextA = 0, vnextB = 0;  // CTA-lifetime virtual block counters (identical in every thread)
    uint32_t it = 0;                  // item parity for the broadcast slots
    uint32_t rp = 0;                  // round parity

    for (;; it ^= 1u) {
        if (tid == 0) {
            unsigned long long item = atomicAdd(work_counter, 1ull);
            misc.item[it] = item;
            if (item < total_items) {
                uint32_t a, b;
                decode_pair(src, item / max_segs, a, b);
                misc.ida[it] = a;
                misc.idb[it] = b;
            }
        }
        __syncthreads();
        const uint64_t item = misc.item[it];
        if (item >= total_items) break;
        const uint64_t pair = item / max_segs;
        const uint32_t seg = (uint32_t)(item % max_segs);
        const SetDesc SA = sets[misc.ida[it]], SB = sets[misc.idb[it]];
        const uint64_t *keysA = use_pal ? SA.pal_keys : SA.keys;
        const uint64_t *keysB = use_pal ? SB.pal_keys : SB.keys;
        const uint32_t nA = use_pal ? SA.n_pal : SA.n;
        const uint32_t nB = use_pal ? SB.n_pal : SB.n;
        const uint64_t L = (uint64_t)nA + nB;
        const uint64_t d0 = (uint64_t)seg * seg_keys;
        if (nA == 0 || nB == 0 || d0 >= L) continue;
        uint64_t d1 = d0 + seg_keys;
        if (d1 > L) d1 = L;
        uint32_t rem = (uint32_t)(d1 - d0);

        uint32_t i0 = 0, j0 = 0;
        if (seg != 0) {
            if (tid < 32) {
                uint32_t s = diag_search_global(keysA, nA, keysB, nB, d0);
                if (lane == 0) misc.start[it] = s;
            }
            __syncthreads();
            i0 = misc.start[it];
            j0 = (uint32_t)(d0 - i0);
        }

        Stream sa, sb;
        stream_begin<C>(sa, keysA, nA, i0, rem, vnextA);
        stream_begin<C>(sb, keysB, nB, j0, rem, vnextB);
        uint32_t cnt = 0;

        while (rem > 0) {
            const uint32_t r = rem < (uint32_t)C::W ? rem : (uint32_t)C::W;
            stream_issue<C>(sa, ringA, barA);
            stream_issue<C>(sb, ringB, barB);
            stream_wait<C>(sa, (sa.head + C::WIN) / ISECT_BLK + 1, barA);
            stream_wait<C>(sb, (sb.head + C::WIN) / ISECT_BLK + 1, barB);
            const uint32_t hA = ringA + sa.hidx * 8u, hB = ringB + sb.hidx * 8u;
            bool is_last;
            uint32_t endA;
            if (r == (uint32_t)C::W) endA = merge_round<C, true>(hA, hB, ringA, ringB, r, cnt, is_last);
            else endA = merge_round<C, false>(hA, hB, ringA, ringB, r, cnt, is_last);
            if (is_last) misc.heads[rp] = endA;
            __syncthreads();  // window fully consumed: heads may move and freed slots may be refilled
            const uint32_t usedA = misc.heads[rp], usedB = r - usedA;
            sa.head += usedA;
            sb.head += usedB;
            sa.hidx += usedA;
            if (sa.hidx >= (uint32_t)C::CAP) sa.hidx -= C::CAP;
            sb.hidx += usedB;
            if (sb.hidx >= (uint32_t)C::CAP) sb.hidx -= C::CAP;
            rem -= r;
            rp ^= 1u;
        }
        // drain copies that were requested but never needed, so the slots can be re-armed
        stream_wait<C>(sa, sa.issued, barA);
        stream_wait<C>(sb, sb.issued, barB);
        vnextA += sa.issued - sa.g0;
        vnextB += sb.issued - sb.g0;

#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0 && cnt) atomicAdd(&counts[pair], cnt);
        __syncthreads();  // every thread has drained before thread 0 re-arms slots for the next item
    }
}

// Configurations tried on B200 (see profiles/): index selected with GKD_ISECT_CFG.  VT is odd on
// purpose: lane l starts about l*VT/2 keys into each ring, and with an even VT that stride maps
// whole half-warps onto 2-4 shared-memory banks (measured 62% of LDS wavefronts were replays).
using Cfg0 = IsectCfg<192, 17, 9, 3>;   // W=3264, 36 KiB rings, 3 CTAs/SM (18 warps)
using Cfg1 = IsectCfg<160, 21, 9, 3>;   // W=3360 (15 warps)
using Cfg2 = IsectCfg<128, 25, 9, 3>;   // W=3200 (12 warps)
using Cfg3 = IsectCfg<224, 15, 9, 3>;   // W=3360 (21 warps)
using Cfg4 = IsectCfg<256, 13, 9, 3>;   // W=3328 (24 warps)
using Cfg5 = IsectCfg<192, 15, 9, 3>;   // W=2880, more prefetch slack
constexpr int N_CFG = 6;
constexpr int DEFAULT_CFG = 0;

static int g_cfg = DEFAULT_CFG;
static int g_algo_mode = 2;  // GKD_ISECT_ALGO: 0 = cta, 1 = warp, 2 = auto (by size balance, see intersect_select)

static int pick_cfg() {
    const char *e = getenv("GKD_ISECT_CFG");
    int c = e ? atoi(e) : DEFAULT_CFG;
    return (c < 0 || c >= N_CFG) ? DEFAULT_CFG : c;
}

#define GKD_FOR_EACH_CFG(X) X(0, Cfg0) X(1, Cfg1) X(2, Cfg2) X(3, Cfg3) X(4, Cfg4) X(5, Cfg5)

cudaError_t intersect_configure() {
    g_cfg = pick_cfg();
    {
        const char *a = getenv("GKD_ISECT_ALGO");
        g_algo_mode = !a ? 2 : (a[0] == 'w' ? 1 : (a[0] == 'c' ? 0 : 2));
    }
    cudaError_t e;
#define X(i, C) \
    if ((e = cudaFuncSetAttribute(k_intersect<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM))) return e;
    GKD_FOR_EACH_CFG(X)
#undef X
    return cudaSuccess;
}

// The warp-cooperative kernel advances both inputs by the keys below min(a[31], b[31]) per step, which
// is efficient when the two sets have similar key density; a pair of very different sizes makes it
// crawl through the denser set 32 keys at a time, where the CTA kernel's merge-path rounds stay
// balanced.  Its fast path also needs the 64 window keys to share their high 32 bits, i.e. sets that
// are dense relative to the key space above bit 32 (42-bit DNA keys of Mbp genomes: yes; 64-bit
// protein keys: no).  Auto mode takes the warp kernel only when both hold; otherwise the CTA kernel.
int intersect_select(uint64_t min_keys, uint64_t max_keys, int key_bits) {
    if (g_algo_mode != 2) return g_algo_mode;
    if (min_keys == 0 || max_keys > 4 * min_keys) return 0;
    const int hi_bits = key_bits > 32 ? key_bits - 32 : 0;
    if (hi_bits >= 40 || (min_keys >> hi_bits) < 1024) return 0;  // fewer than ~1000 keys per high-word value
    return 1;
}

int intersect_items_per_sm(int algo) { return algo ? WK_CTAS * WK_WARPS * 6 : 3 * 8; }

int intersect_min_segment(int algo) {
    if (algo == 1) return 1024;
#define X(i, C) \
    if (g_cfg == i) return C::W;
    GKD_FOR_EACH_CFG(X)
#undef X
    return ISECT_W_MAX;
}

template <class C>
static cudaError_t launch_cfg(const SetDesc *sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                              uint32_t *counts, unsigned long long *work_counter, int n_sms, cudaStream_t s) {
    uint64_t items = src.count * (uint64_t)max_segs;
    uint64_t grid = (uint64_t)n_sms * C::CTAS;  // persistent: every SM holds CTAS resident CTAs
    if (grid > items) grid = items;
    k_intersect<C><<<(unsigned)grid, C::THREADS, C::SMEM, s>>>(sets, src, use_pal, seg_keys, max_segs, counts, work_counter);
    return cudaGetLastError();
}

static cudaError_t launch_warp(const SetDesc *sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                               uint32_t *counts, unsigned long long *work_counter, int n_sms, cudaStream_t s);

cudaError_t launch_intersect(const SetDesc *sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                             uint32_t *counts, unsigned long long *work_counter, int n_sms, int algo, cudaStream_t s) {
    if (src.count == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (algo == 1) return launch_warp(sets, src, use_pal, seg_keys, max_segs, counts, work_counter, n_sms, s);
#define X(i, C) \
    if (g_cfg == i) return launch_cfg<C>(sets, src, use_pal, seg_keys, max_segs, counts, work_counter, n_sms, s);
    GKD_FOR_EACH_CFG(X)
#undef X
    return cudaErrorInvalidValue;
}

// ---- kernel 4, warp-cooperative variant ----------------------------------------------------------------
// One warp owns a work item and streams both inputs through its own pair of 4 KiB rings (TMA bulk
// copies of 128 keys, one mbarrier per slot).  A step loads 32 consecutive keys of each input with
// coalesced, bank-conflict-free LDS.64 (one key per lane), takes x = min(a[31], b[31]), consumes every
// key <= x of both windows (__ballot_sync + __popc give the two advances), and counts matches by a
// 5-round shuffle binary search of each lane's A key in the B window.  No per-thread partition search,
// no random shared-memory access, no CTA barrier in the streaming loop.
constexpr int WK_BLK = GKD_WK_BLK;           // keys per TMA block (2 KiB: fewer, larger copies beat 4 x 1 KiB, the mbarrier traffic being the cost)
#ifndef GKD_WK_NBLK
#define GKD_WK_NBLK 2
#endif
constexpr int WK_NBLK = GKD_WK_NBLK;          // ring slots per input
constexpr int WK_BATCH = WK_NBLK / 2;          // slots that must be free before more copies are requested
constexpr int WK_CAP = WK_BLK * WK_NBLK;     // 512 keys = 4 KiB per input ring
constexpr uint32_t WK_SMEM = WK_WARPS * (2 * WK_CAP * 8 + 2 * WK_NBLK * 8);

struct WStream {
    const uint64_t *keys;
    uint32_t pos;     // next unconsumed key
    uint32_t shift;   // ring index of position p is (p + shift) & (WK_CAP - 1)
    uint32_t g0, v0;  // first block of the item and its warp-lifetime virtual block number
    uint32_t issued, ready, limit;  // block numbers: next to request / first not known landed / one past last
};

__device__ __forceinline__ void wk_begin(WStream &s, const uint64_t *keys, uint32_t n, uint32_t pos0, uint32_t last_pos,
                                         uint32_t vnext) {
    s.keys = keys;
    s.pos = pos0;
    s.g0 = pos0 / WK_BLK;
    s.v0 = vnext;
    s.shift = (vnext - s.g0) * WK_BLK;
    s.issued = s.g0;
    s.ready = s.g0;
    if (last_pos > n) last_pos = n;
    s.limit = (last_pos + 32) / WK_BLK + 1;  // a step reads [pos, pos + 32)
}

__device__ __forceinline__ void wk_issue(WStream &s, uint32_t ring_addr, uint32_t bar_addr, uint32_t lane) {
    uint32_t upto = s.pos / WK_BLK + WK_NBLK;
    if (upto > s.limit) upto = s.limit;
    if (upto > s.issued) {
        __syncwarp();  // every lane is done reading the slots that are about to be overwritten
        if (lane == 0) {
            for (uint32_t g = s.issued; g < upto; g++) {
                const uint32_t slot = (s.v0 + (g - s.g0)) & (WK_NBLK - 1);
                mbar_expect_tx(bar_addr + slot * 8, WK_BLK * 8);
                tma_load_1d(ring_addr + slot * (WK_BLK * 8), s.keys + (size_t)g * WK_BLK, WK_BLK * 8, bar_addr + slot * 8);
            }
        }
        s.issued = upto;
    }
}

__device__ __forceinline__ void wk_wait(WStream &s, uint32_t upto, uint32_t bar_addr) {
    while (s.ready < upto) {
        const uint32_t v = s.v0 + (s.ready - s.g0);
        mbar_wait(bar_addr + (v & (WK_NBLK - 1)) * 8, (v / WK_NBLK) & 1u);
        s.ready++;
    }
}

__device__ __forceinline__ void wk_poll(WStream &s, uint32_t bar_addr);

// one bookkeeping visit for a stream; returns the position at which the next visit is due
__device__ __forceinline__ uint32_t wk_service(WStream &s, uint32_t ring_addr, uint32_t bar_addr, uint32_t lane) {
    wk_issue(s, ring_addr, bar_addr, lane);
    const uint32_t need = (s.pos + 31) / WK_BLK + 1;
    wk_wait(s, need < s.limit ? need : s.limit, bar_addr);
    if (WK_NBLK > 2) wk_poll(s, bar_addr);
    const uint32_t ti = s.issued < s.limit ? (s.issued + WK_BATCH - WK_NBLK) * WK_BLK : 0xFFFFFFFFu;
    const uint32_t tw = s.ready < s.limit ? s.ready * WK_BLK - 31 : 0xFFFFFFFFu;
    return ti < tw ? ti : tw;
}

// advance `ready` over blocks that have already landed, without blocking; every lane must have seen
// the phase complete (that observation is what makes the copied bytes visible to it)
__device__ __forceinline__ void wk_poll(WStream &s, uint32_t bar_addr) {
    while (s.ready < s.issued) {
        const uint32_t v = s.v0 + (s.ready - s.g0);
        const bool ok = mbar_test(bar_addr + (v & (WK_NBLK - 1)) * 8, (v / WK_NBLK) & 1u);
        if (!__all_sync(0xffffffffu, ok)) break;
        s.ready++;
    }
}

__global__ void __launch_bounds__(WK_WARPS * 32, WK_CTAS)
    k_intersect_warp(const SetDesc *__restrict__ sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                     uint32_t *__restrict__ counts, unsigned long long *__restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = smem_u32(smem_raw) + warp * (2 * WK_CAP * 8);
    const uint32_t ringA = base, ringB = base + WK_CAP * 8;
    const uint32_t barA = smem_u32(smem_raw) + WK_WARPS * (2 * WK_CAP * 8) + warp * (2 * WK_NBLK * 8);
    const uint32_t barB = barA + WK_NBLK * 8;
    if (lane == 0) {
        for (int i = 0; i < WK_NBLK; i++) {
            mbar_init(barA + i * 8, 1);
            mbar_init(barB + i * 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint64_t total_items = src.count * (uint64_t)max_segs;
    uint32_t vnextA = 0, vnextB = 0;
    for (;;) {
        unsigned long long item = 0;
        uint32_t ida = 0, idb = 0;
        if (lane == 0) {
            item = atomicAdd(work_counter, 1ull);
            if (item < total_items) decode_pair(src, item / max_segs, ida, idb);
        }
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total_items) break;
        ida = __shfl_sync(0xffffffffu, ida, 0);
        idb = __shfl_sync(0xffffffffu, idb, 0);
        const uint64_t pair = item / max_segs;
        const uint32_t seg = (uint32_t)(item % max_segs);
        const SetDesc SA = sets[ida], SB = sets[idb];
        const uint64_t *keysA = use_pal ? SA.pal_keys : SA.keys;
        const uint64_t *keysB = use_pal ? SB.pal_keys : SB.keys;
        const uint32_t nA = use_pal ? SA.n_pal : SA.n;
        const uint32_t nB = use_pal ? SB.n_pal : SB.n;
        const uint64_t L = (uint64_t)nA + nB;
        const uint64_t d0 = (uint64_t)seg * seg_keys;
        if (nA == 0 || nB == 0 || d0 >= L) continue;
        uint64_t d1 = d0 + seg_keys;
        if (d1 > L) d1 = L;
        // segment = merge-path diagonals [d0, d1): A keys [i0, i1) are mine; B is read from j0 on
        const uint32_t i0 = d0 ? diag_search_global(keysA, nA, keysB, nB, d0) : 0u;
        const uint32_t j0 = (uint32_t)(d0 - i0);
        const uint32_t i1 = d1 < L ? diag_search_global(keysA, nA, keysB, nB, d1) : nA;
        const uint32_t j1 = (uint32_t)(d1 - i1);

        WStream sa, sb;
        wk_begin(sa, keysA, nA, i0, i1, vnextA);
        wk_begin(sb, keysB, nB, j0, j1 + 1, vnextB);  // B[j1] may equal my last A key
        uint32_t cnt = 0;
        const uint32_t loadedEndB = sb.limit * WK_BLK;  // B keys at or past this position are never loaded
        // steps whose 32-key windows lie inside [.., i1) and the loaded part of B take the interior path
        const int edgeA = (int)i1 - 32, edgeB = (int)loadedEndB - 32;
        uint32_t trigA = 0, trigB = 0;  // ring bookkeeping is due once pos reaches these positions
        // per-lane byte offset of this lane's key inside each ring, advanced by the consumed counts
        constexpr uint32_t RING_MASK = (WK_CAP - 1) * 8u;
        uint32_t laneA = ((sa.pos + sa.shift + lane) * 8u) & RING_MASK, laneB = ((sb.pos + sb.shift + lane) * 8u) & RING_MASK;
        while (sa.pos < i1) {  // only my A keys can produce matches; leftover B keys need no visit
            // ring bookkeeping only when a stream reaches its next trigger position: request the blocks
            // whose slots are free, make sure the window [pos, pos + 32) has landed, and compute the next
            // position at which to look again (a slot frees up, or the landed data runs out)
            if (sa.pos >= trigA) trigA = wk_service(sa, ringA, barA, lane);
            if (sb.pos >= trigB) trigB = wk_service(sb, ringB, barB, lane);
            const uint32_t pa = sa.pos + lane, pb = sb.pos + lane;
            const uint32_t adrA = ringA + laneA;
            const uint32_t adrB = ringB + laneB;
            uint32_t nAc, nBc;
            if ((int)sa.pos <= edgeA && (int)sb.pos <= edgeB) {
                // interior step: every lane holds a loaded key of each input
                const uint64_t a = lds64(adrA);
                const uint64_t b = lds64(adrB);
                const uint32_t ah = (uint32_t)(a >> 32), al = (uint32_t)a, bh = (uint32_t)(b >> 32), bl = (uint32_t)b;
                const uint32_t href = __shfl_sync(0xffffffffu, ah, 0);
                if (__all_sync(0xffffffffu, ah == href && bh == href)) {
                    // all 64 keys share their high word (true for ~99 % of the windows of 42-bit keys):
                    // the whole step runs on the low words
                    const uint32_t amax = __shfl_sync(0xffffffffu, al, 31), bmax = __shfl_sync(0xffffffffu, bl, 31);
                    const uint32_t x = amax < bmax ? amax : bmax;
                    const bool ca = al <= x;
                    nAc = __popc(__ballot_sync(0xffffffffu, ca));
                    nBc = __popc(__ballot_sync(0xffffffffu, bl <= x));
                    // lower bound of al among the 32 sorted low words of B: 4-ary, 4-ary, binary
                    // (7 shuffles in 3 dependent levels instead of 5 dependent ones)
                    const uint32_t q1 = __shfl_sync(0xffffffffu, bl, 7), q2 = __shfl_sync(0xffffffffu, bl, 15),
                                   q3 = __shfl_sync(0xffffffffu, bl, 23);
                    uint32_t lo = q3 < al ? 24u : (q2 < al ? 16u : (q1 < al ? 8u : 0u));  // probes are sorted
                    const uint32_t r1 = __shfl_sync(0xffffffffu, bl, lo + 1), r2 = __shfl_sync(0xffffffffu, bl, lo + 3),
                                   r3 = __shfl_sync(0xffffffffu, bl, lo + 5);
                    lo += r3 < al ? 6u : (r2 < al ? 4u : (r1 < al ? 2u : 0u));
                    const uint32_t t0 = __shfl_sync(0xffffffffu, bl, lo);
                    const uint32_t t1 = __shfl_sync(0xffffffffu, bl, lo + 1 > 31u ? 31u : lo + 1);
                    // the key equals B[lo] or, if B[lo] is smaller, B[lo+1]
                    cnt += __popc(__ballot_sync(0xffffffffu, ca && (t0 == al || t1 == al)));
                } else {
                    const uint64_t amax = __shfl_sync(0xffffffffu, a, 31), bmax = __shfl_sync(0xffffffffu, b, 31);
                    const uint64_t x = amax < bmax ? amax : bmax;
                    const bool ca = a <= x;
                    nAc = __popc(__ballot_sync(0xffffffffu, ca));
                    nBc = __popc(__ballot_sync(0xffffffffu, b <= x));
                    uint32_t lo = 0;
#pragma unroll
                    for (int st = 16; st >= 1; st >>= 1) {
                        const uint64_t bv = __shfl_sync(0xffffffffu, b, lo + st - 1);
                        if (bv < a) lo += st;
                    }
                    const uint64_t bm = __shfl_sync(0xffffffffu, b, lo);
                    cnt += __popc(__ballot_sync(0xffffffffu, ca && bm == a));
                }
            } else {
                // edge step: keys past my A range or past the loaded B blocks read as sentinels
                const bool va = pa < i1;
                const bool vb = pb < loadedEndB;
                uint64_t a = KEY_SENTINEL, b = KEY_SENTINEL;
                if (va) a = lds64(adrA);
                if (vb) b = lds64(adrB);
                const uint64_t amax = __shfl_sync(0xffffffffu, a, 31), bmax = __shfl_sync(0xffffffffu, b, 31);
                const uint64_t x = amax < bmax ? amax : bmax;
                const bool ca = va && a <= x;
                nAc = __popc(__ballot_sync(0xffffffffu, ca));
                nBc = __popc(__ballot_sync(0xffffffffu, b <= x));
                uint32_t lo = 0;
#pragma unroll
                for (int st = 16; st >= 1; st >>= 1) {
                    const uint64_t bv = __shfl_sync(0xffffffffu, b, lo + st - 1);
                    if (bv < a) lo += st;
                }
                const uint64_t bm = __shfl_sync(0xffffffffu, b, lo);
                cnt += __popc(__ballot_sync(0xffffffffu, ca && bm == a));
            }
            sa.pos += nAc;
            sb.pos += nBc;
            laneA = (laneA + nAc * 8u) & RING_MASK;
            laneB = (laneB + nBc * 8u) & RING_MASK;
        }
        // drain copies that were requested but never needed, so the slots can be re-armed
        wk_wait(sa, sa.issued, barA);
        wk_wait(sb, sb.issued, barB);
        vnextA += sa.issued - sa.g0;
        vnextB += sb.issued - sb.g0;
        if (lane == 0 && cnt) atomicAdd(&counts[pair], cnt);
    }
}

static cudaError_t launch_warp(const SetDesc *sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                               uint32_t *counts, unsigned long long *work_counter, int n_sms, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_intersect_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WK_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    uint64_t items = src.count * (uint64_t)max_segs;
    uint64_t grid = (uint64_t)n_sms * WK_CTAS;
    uint64_t need = (items + WK_WARPS - 1) / WK_WARPS;
    if (grid > need) grid = need;
    k_intersect_warp<<<(unsigned)grid, WK_WARPS * 32, WK_SMEM, s>>>(sets, src, use_pal, seg_keys, max_segs, counts, work_counter);
    return cudaGetLastError();
}

// ---- kernel 4, small-set variant ---------------------------------------------------------------------
// When every set is small (gene/protein-sized FASTA records: the whole pair is a few KiB and stays in
// L1/L2) the streaming kernel's per-item set-up dominates.  Here one warp owns a pair: each lane takes
// keys of the smaller set and binary-searches them in the larger one (sentinel-padded, so no bound
// check on the final probe); counts are combined with shuffles and written without atomics.
constexpr uint32_t SMALL_SET_MAX_KEYS = 16384;

__global__ void __launch_bounds__(256)
    k_intersect_small(const SetDesc *__restrict__ sets, PairSource src, int use_pal, uint32_t *__restrict__ counts) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t t = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < src.count; t += warps) {
        uint32_t ida, idb;
        decode_pair(src, t, ida, idb);
        const SetDesc SA = sets[ida], SB = sets[idb];
        const uint64_t *ka = use_pal ? SA.pal_keys : SA.keys, *kb = use_pal ? SB.pal_keys : SB.keys;
        uint32_t na = use_pal ? SA.n_pal : SA.n, nb = use_pal ? SB.n_pal : SB.n;
        if (na > nb) {
            const uint64_t *tk = ka;
            ka = kb;
            kb = tk;
            uint32_t tn = na;
            na = nb;
            nb = tn;
        }
        uint32_t cnt = 0;
        if (na != 0) {  // then nb != 0 and kb is padded with sentinels past nb
            for (uint32_t i = lane; i < na; i += 32) {
                const uint64_t key = __ldg(ka + i);
                uint32_t lo = 0, hi = nb;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (__ldg(kb + mid) < key) lo = mid + 1;
                    else hi = mid;
                }
                cnt += (__ldg(kb + lo) == key) ? 1u : 0u;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) counts[t] = cnt;
    }
}

cudaError_t launch_intersect_small(const SetDesc *sets, PairSource src, int use_pal, uint32_t *counts, int n_sms,
                                   cudaStream_t s) {
    if (src.count == 0) return cudaSuccess;
    uint64_t grid = (src.count + 7) / 8;
    if (grid > (uint64_t)n_sms * 8) grid = (uint64_t)n_sms * 8;
    k_intersect_small<<<(unsigned)grid, 256, 0, s>>>(sets, src, use_pal, counts);
    return cudaGetLastError();
}

uint32_t intersect_small_max_keys() { return SMALL_SET_MAX_KEYS; }

// ---- kernel 5: distance epilogue ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_epilogue(const SetDesc *__restrict__ sets, PairSource src, const uint32_t *__restrict__ counts,
               const uint32_t *__restrict__ pal_counts, int both_strands, uint64_t *__restrict__ inter,
               double *__restrict__ dist) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= src.count) return;
    uint32_t ida, idb;
    decode_pair(src, t, ida, idb);
    const SetDesc A = sets[ida], B = sets[idb];
    uint64_t c = counts[t];
    uint64_t I, sa, sb;
    if (both_strands) {
        // the reference's sets hold both strands: |S| = 2|C| - P, I = 2|C_A n C_B| - P(C_A n C_B)
        uint64_t cp = pal_counts ? pal_counts[t] : 0;
        I = 2 * c - cp;
        sa = 2ull * A.n - A.n_pal;
        sb = 2ull * B.n - B.n_pal;
    } else {
        I = c;
        sa = A.n;
        sb = B.n;
    }
    if (inter) inter[t] = I;
    if (dist) {
        double ret = 1.0;
        double similarity = (double)I;
        if (similarity > 0) {
            // (this.size() + other.size()) is a Java int addition
            int32_t sum = (int32_t)((uint32_t)sa + (uint32_t)sb);
            double uni = (double)sum - similarity;
            ret = 1.0 - similarity / uni;
        }
        dist[t] = ret;
    }
}

cudaError_t launch_epilogue(const SetDesc *sets, PairSource src, const uint32_t *counts, const uint32_t *pal_counts,
                            int both_strands, uint64_t *inter, double *dist, cudaStream_t s) {
    if (src.count == 0) return cudaSuccess;
    uint64_t blocks = (src.count + 255) / 256;
    k_epilogue<<<(unsigned)blocks, 256, 0, s>>>(sets, src, counts, pal_counts, both_strands, inter, dist);
    return cudaGetLastError();
}

}  // namespace gkd
