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
//   * persistent CTAs (3 per SM, 256 threads) pull work items (pair, merge-path segment) from one
//     global counter; pairs are enumerated row-major so CTAs resident at the same time mostly share
//     the row genome and it is served from the 126 MB L2 instead of HBM;
//   * each input is streamed through a 32 KiB shared-memory ring filled by TMA bulk copies
//     (cp.async.bulk, 4 KiB blocks, one mbarrier per ring slot); sets carry a sentinel tail so no
//     bounds checks are needed; the ring is refilled as soon as a block is consumed, which keeps
//     about two rounds of loads in flight per CTA;
//   * a round merges W = 2048 keys: every thread finds its merge-path split in shared memory and
//     then merges 8 keys serially, counting equal heads; the last thread's end point advances the
//     stream heads;
//   * segment starts inside a pair are found by a warp-cooperative 32-ary merge-path search on
//     global memory (__ballot_sync / __popc select the sub-range);
//   * per-warp shuffle reduction, one atomicAdd per warp per item.
// Bound: HBM (or L2 when the row set is resident).  Algorithmic bytes: 8 * (|A| + |B|) per pair.
#include "gkd_internal.cuh"

namespace gkd {

// ---- mbarrier / TMA wrappers ------------------------------------------------------------------------
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
    uint32_t done = 0;
    uint32_t spins = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 24)) __trap();  // a lost copy must not hang the GPU
    }
}

struct __align__(128) IsectSmem {
    uint64_t ring[2][ISECT_CAP];
    uint64_t bar[2][ISECT_NBLK];
    unsigned long long item[2];
    uint32_t ida[2], idb[2];
    uint32_t heads[2][2];
    uint32_t start[2];
};

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

// One input stream: a ring of ISECT_NBLK blocks of a sentinel-padded sorted key array.
struct Stream {
    const uint64_t *keys;  // global base of the set
    uint32_t head;         // next unconsumed key
    uint32_t g0;           // first block of this item
    uint32_t v0;           // virtual (CTA-lifetime) block counter at g0: slot = v % NBLK, parity = (v / NBLK) & 1
    uint32_t issued;       // next block to request
    uint32_t ready;        // blocks < ready are known to have landed
    uint32_t limit;        // one past the last block this item can touch
    uint32_t shift;        // ring index of key position p is (p + shift) & (CAP - 1)
};

__device__ __forceinline__ void stream_begin(Stream &s, const uint64_t *keys, uint32_t n, uint32_t head0,
                                             uint32_t max_consume, uint32_t vnext) {
    s.keys = keys;
    s.head = head0;
    s.g0 = head0 / ISECT_BLK;
    s.v0 = vnext;
    s.issued = s.g0;
    s.ready = s.g0;
    uint64_t last_pos = (uint64_t)head0 + max_consume;
    if (last_pos > n) last_pos = n;
    s.limit = (uint32_t)((last_pos + ISECT_W) / ISECT_BLK) + 1;
    s.shift = (s.v0 - s.g0) * (uint32_t)ISECT_BLK;
}

// request every block whose ring slot is free (thread 0 issues; all threads track the counter)
__device__ __forceinline__ void stream_issue(Stream &s, uint32_t ring_addr, uint32_t bar_addr) {
    uint32_t upto = s.head / ISECT_BLK + ISECT_NBLK;
    if (upto > s.limit) upto = s.limit;
    if (threadIdx.x == 0) {
        for (uint32_t g = s.issued; g < upto; g++) {
            uint32_t slot = (s.v0 + (g - s.g0)) % ISECT_NBLK;
            uint32_t bar = bar_addr + slot * 8;
            mbar_expect_tx(bar, ISECT_BLK * 8);
            tma_load_1d(ring_addr + slot * (ISECT_BLK * 8), s.keys + (size_t)g * ISECT_BLK, ISECT_BLK * 8, bar);
        }
    }
    if (upto > s.issued) s.issued = upto;
}

// block until blocks [ready, upto) have landed
__device__ __forceinline__ void stream_wait(Stream &s, uint32_t upto, uint32_t bar_addr) {
    for (uint32_t g = s.ready; g < upto; g++) {
        uint32_t v = s.v0 + (g - s.g0);
        mbar_wait(bar_addr + (v % ISECT_NBLK) * 8, (v / ISECT_NBLK) & 1u);
    }
    if (upto > s.ready) s.ready = upto;
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

__global__ void __launch_bounds__(ISECT_THREADS, 3)
    k_intersect(const SetDesc *__restrict__ sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                uint32_t *__restrict__ counts, unsigned long long *__restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    IsectSmem &sm = *reinterpret_cast<IsectSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const uint32_t ringA = smem_u32(&sm.ring[0][0]), ringB = smem_u32(&sm.ring[1][0]);
    const uint32_t barA = smem_u32(&sm.bar[0][0]), barB = smem_u32(&sm.bar[1][0]);
    constexpr uint32_t M = ISECT_CAP - 1;

    if (tid == 0) {
        for (int s = 0; s < ISECT_NBLK; s++) {
            mbar_init(barA + s * 8, 1);
            mbar_init(barB + s * 8, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint64_t total_items = src.count * (uint64_t)max_segs;
    uint32_t vnextA = 0, vnextB = 0;  // CTA-lifetime virtual block counters (identical in every thread)
    uint32_t it = 0;                  // item parity for the broadcast slots
    uint32_t round_parity = 0;

    for (;; it ^= 1u) {
        if (tid == 0) {
            unsigned long long item = atomicAdd(work_counter, 1ull);
            sm.item[it] = item;
            if (item < total_items) {
                uint32_t a, b;
                decode_pair(src, item / max_segs, a, b);
                sm.ida[it] = a;
                sm.idb[it] = b;
            }
        }
        __syncthreads();
        const uint64_t item = sm.item[it];
        if (item >= total_items) break;
        const uint64_t pair = item / max_segs;
        const uint32_t seg = (uint32_t)(item % max_segs);
        const SetDesc SA = sets[sm.ida[it]], SB = sets[sm.idb[it]];
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
                if (lane == 0) sm.start[it] = s;
            }
            __syncthreads();
            i0 = sm.start[it];
            j0 = (uint32_t)(d0 - i0);
        }

        Stream sa, sb;
        stream_begin(sa, keysA, nA, i0, rem, vnextA);
        stream_begin(sb, keysB, nB, j0, rem, vnextB);
        uint32_t cnt = 0;

        while (rem > 0) {
            const uint32_t r = rem < (uint32_t)ISECT_W ? rem : (uint32_t)ISECT_W;
            stream_issue(sa, ringA, barA);
            stream_issue(sb, ringB, barB);
            {
                uint32_t needA = (sa.head + ISECT_W) / ISECT_BLK + 1;
                uint32_t needB = (sb.head + ISECT_W) / ISECT_BLK + 1;
                stream_wait(sa, needA < sa.limit ? needA : sa.limit, barA);
                stream_wait(sb, needB < sb.limit ? needB : sb.limit, barB);
            }
            const uint32_t baseA = sa.head + sa.shift, baseB = sb.head + sb.shift;
            // merge-path split of this thread's diagonal inside the window (A first on ties)
            uint32_t d = (uint32_t)tid * ISECT_VT;
            if (d > r) d = r;
            uint32_t lo = 0, hi = d;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                uint64_t a = sm.ring[0][(baseA + mid) & M];
                uint64_t b = sm.ring[1][(baseB + d - 1 - mid) & M];
                if (a <= b) lo = mid + 1;
                else hi = mid;
            }
            uint32_t ia = lo, ib = d - lo;
            uint32_t steps = r - d;
            if (steps > (uint32_t)ISECT_VT) steps = ISECT_VT;
            uint32_t pa = (baseA + ia) & M, pb = (baseB + ib) & M;
            uint64_t a = sm.ring[0][pa], b = sm.ring[1][pb];
#pragma unroll
            for (int s = 0; s < ISECT_VT; s++) {
                if ((uint32_t)s < steps) {
                    bool take_a = a <= b;
                    cnt += (a == b) ? 1u : 0u;  // counted once, when the A copy is consumed
                    if (take_a) {
                        pa = (pa + 1) & M;
                        ia++;
                        a = sm.ring[0][pa];
                    } else {
                        pb = (pb + 1) & M;
                        ib++;
                        b = sm.ring[1][pb];
                    }
                }
            }
            if (d < r && d + ISECT_VT >= r) {  // this thread ends exactly on the round's last diagonal
                sm.heads[round_parity][0] = ia;
                sm.heads[round_parity][1] = ib;
            }
            __syncthreads();
            sa.head += sm.heads[round_parity][0];
            sb.head += sm.heads[round_parity][1];
            rem -= r;
            round_parity ^= 1u;
        }
        // drain copies that were requested but never needed, so the slots can be re-armed
        stream_wait(sa, sa.issued, barA);
        stream_wait(sb, sb.issued, barB);
        vnextA += sa.issued - sa.g0;
        vnextB += sb.issued - sb.g0;

#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0 && cnt) atomicAdd(&counts[pair], cnt);
        __syncthreads();  // every thread is done with the ring before the next item re-arms it
    }
}

static int g_isect_smem = 0;

cudaError_t intersect_configure() {
    g_isect_smem = (int)sizeof(IsectSmem);
    return cudaFuncSetAttribute(k_intersect, cudaFuncAttributeMaxDynamicSharedMemorySize, g_isect_smem);
}

cudaError_t launch_intersect(const SetDesc *sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                             uint32_t *counts, unsigned long long *work_counter, int n_sms, cudaStream_t s) {
    if (src.count == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    uint64_t items = src.count * (uint64_t)max_segs;
    uint64_t grid = (uint64_t)n_sms * 3;  // persistent: 3 resident CTAs per SM
    if (grid > items) grid = items;
    k_intersect<<<(unsigned)grid, ISECT_THREADS, g_isect_smem, s>>>(sets, src, use_pal, seg_keys, max_segs, counts,
                                                                    work_counter);
    return cudaGetLastError();
}

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
