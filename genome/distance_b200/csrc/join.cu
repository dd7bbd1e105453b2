// join.cu -- kernel 4, block-join form: a block of 64 (or 32) row sets against every column set at once.
//
// Reference semantics restated (same as intersect.cu): SequenceKmers.similarity = number of members of one
// HashSet<String> found in the other, for every pair the caller enumerates -- the strict upper triangle of
// FastaDistanceProcessor.java:174-194 or the query x reference rectangle of GenomeProcessor.java:129-147.
// The reference probes one hash set with the members of the other, pair by pair.  The bucket-merge kernel
// (intersect.cu) streams both sets of every pair; for a pair MATRIX that repeats the same work row after row:
// a column key is compared with every row separately.  This kernel keeps the reference's probe formulation but
// shares the probe between the rows of a block:
//
//   * the key space is cut into 2^L equal ranges (L per class of row sizes, chosen so that the keys a block of
//     rows holds in one range fill ~30 % of the table).  Sets are stored in mixed-key order with bucket offset
//     tables (gkd_internal.cuh), so the keys any set holds in a range are one contiguous run of its low words;
//   * a task = (row block, range).  The CTA builds ONE open-addressing table in shared memory: key (32-bit low
//     word -- the range pins the other bits -- or the whole 64-bit h of wide keys) -> one 32-bit mask word per 32
//     rows naming the rows that hold it (find-or-insert with atomicCAS on the key word, atomicOr on the mask word:
//     lock-free, no ordering between rows needed; the runs' total is checked first, so the table cannot fill up);
//   * then every column set's run of the same range is streamed once from L2/HBM (coalesced, KPT keys per lane,
//     the next column's keys requested before the current one is probed) and probed; a hit returns the row mask.
//     Misses -- nearly every probe of an unrelated genome -- settle all rows of the block at once;
//   * hits are tallied in packed byte counters per lane, summed over the warp with a reduce-scatter on the packed
//     words for columns that had a match, and flushed with one atomicAdd per (row, column, range) that saw one;
//   * warp 0 looks up the next task (rows, runs) while the other warps still probe the current table.
//
// Per pair and column key this is 1/64 (1/32) of a probe instead of a two-pointer merge step, and every set is read
// once per ROW BLOCK instead of once per row; tasks are enumerated range-major so the column runs of one range are
// served from L2 to all row blocks (HBM sees every set once per call).  Counts are exact (a key's identity inside a
// range is its low word).  Serves 32-bit low words (DNA/RNA K <= 21, protein K <= 5) and 64-bit keys when a call has
// enough rows and columns; everything else (lists, greedy pass, tiny sets, thin slices, palindrome sub-sets) stays on
// the merge kernel.  DESIGN.md section 4 has the cost model and the measurements.
#include <cstdlib>

#include "gkd_internal.cuh"

namespace gkd {

namespace {

constexpr uint32_t JOIN_INVALID = 0xFFFFFFFFu;

// where set S keeps the keys of range rho (level L): run [lo, hi) of its low words, plus a value filter
// when the set's own table is coarser than the range (fm == 0: every key of the run is in the range)
template <typename KT>
struct RangeRun {
    const KT *lows;
    uint32_t lo, hi;
    KT fm, fv;
};

template <typename KT>
__device__ __forceinline__ RangeRun<KT> range_run(const SubSet &S, uint32_t L, uint32_t rho) {
    RangeRun<KT> r;
    r.lows = (const KT *)S.lows;
    r.lo = r.hi = 0;
    r.fm = r.fv = 0;
    if (S.n == 0) return r;
    if (S.level >= L) {
        const uint32_t sh = S.level - L;
        r.lo = __ldg(S.offs + ((size_t)rho << sh));
        r.hi = __ldg(S.offs + ((size_t)(rho + 1) << sh));
    } else {  // coarser table: take the enclosing bucket and keep the keys whose next bits select this range
        const uint32_t d = L - S.level;
        const uint32_t c = rho >> d;
        r.lo = __ldg(S.offs + c);
        r.hi = __ldg(S.offs + c + 1);
        r.fm = (KT)(((unsigned long long)1 << d) - 1ull);
        r.fv = (KT)rho & r.fm;
    }
    return r;
}

// slot * 4 of a key (the byte offset of its mask word; the key word is at the same offset for 32-bit keys, twice
// that for 64-bit keys)
template <int SLOTS_LOG2>
__device__ __forceinline__ uint32_t join_off4(uint32_t k) {
    return ((k * 0x9E3779B1u) >> (30 - SLOTS_LOG2)) & (((1u << SLOTS_LOG2) - 1u) << 2);
}
template <int SLOTS_LOG2>
__device__ __forceinline__ uint32_t join_off4(uint64_t k) {
    return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> (62 - SLOTS_LOG2)) & (((1u << SLOTS_LOG2) - 1u) << 2);
}

// shared-window addresses: no generic-pointer set-up per load
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_key(uint32_t kbase, uint32_t off4, uint32_t) { return lds_u32(kbase + off4); }
__device__ __forceinline__ uint64_t lds_key(uint32_t kbase, uint32_t off4, uint64_t) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(kbase + 2u * off4));
    return v;
}
__device__ __forceinline__ uint32_t cas_key(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ uint64_t cas_key(uint64_t *p, uint64_t cmp, uint64_t v) {
    return atomicCAS((unsigned long long *)p, (unsigned long long)cmp, (unsigned long long)v);
}
__device__ __forceinline__ uint32_t ldg_key(const uint32_t *p) { return __ldg(p); }
__device__ __forceinline__ uint64_t ldg_key(const uint64_t *p) { return __ldg((const unsigned long long *)p); }

}  // namespace

// KT = key word of the sets (uint32_t low words, or the full uint64_t h of keys wider than 42 bits), ROWS = rows per
// block (32 or 64: one or two mask words per slot), KPT = column keys per lane per trip (a column's run of one range,
// fill / ROWS keys on average, should fit one trip)
template <typename KT, int SLOTS_LOG2, int THREADS, int CTAS, int ROWS, int KPT>
__global__ void __launch_bounds__(THREADS, CTAS)
    k_join(const SetDesc *__restrict__ sets, JoinPlan plan, uint32_t *__restrict__ counts,
           unsigned long long *__restrict__ work_counter, uint32_t *__restrict__ err) {
    constexpr uint32_t SLOTS = 1u << SLOTS_LOG2, SMASK = SLOTS - 1u;
    constexpr uint32_t MAXFILL = SLOTS / 2u + SLOTS / 8u;
    constexpr int NW = THREADS / 32, RW = ROWS / 32;
    static_assert(ROWS == 32 || ROWS == 64, "one or two mask words per slot");
    static_assert(THREADS >= ROWS, "one thread per row looks the runs up");
    extern __shared__ __align__(16) uint32_t sm_tab[];
    constexpr uint32_t KS = (uint32_t)sizeof(KT);
    KT *keys = (KT *)sm_tab;
    uint32_t *masks = sm_tab + SLOTS * (KS / 4u);  // masks: word w of slot s at masks[w * SLOTS + s]
    // Everything a task needs before its table can be built, looked up by warp 0 WHILE the other warps still probe
    // the previous task's table (the look-ups are a chain of dependent global loads): double-buffered.
    struct TaskInfo {
        unsigned long long task;
        uint32_t level, rho, fill, min_id;
        uint32_t rowid[ROWS], rowpos[ROWS];
        RangeRun<KT> run[ROWS];
    };
    __shared__ TaskInfo s_ti[2];
    __shared__ uint32_t s_colgrp;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t FULL = 0xffffffffu;
    // shared-window address of the table, made opaque so it stays in a register instead of being re-derived
    // from the generic pointer at every probe
    uint32_t kbase = (uint32_t)__cvta_generic_to_shared(keys);
    asm volatile("" : "+r"(kbase));
    constexpr uint32_t M4 = SMASK << 2, MASKS_OFF = SLOTS * KS;  // byte offset of mask word 0 behind the keys

    // warp 0: take the next task from the global counter and look up its rows and their runs
    auto prepare = [&](TaskInfo &ti) {
        unsigned long long task = 0;
        if (lane == 0) task = atomicAdd(work_counter, 1ull);
        task = __shfl_sync(FULL, task, 0);
        if (lane == 0) ti.task = task;
        if (task >= plan.n_tasks) return;
        // task -> (class, range, row block); range-major inside a class
        uint32_t ci = 0;
        while (ci + 1 < plan.n_classes && task >= plan.cls[ci + 1].task_first) ci++;
        const uint32_t L = plan.cls[ci].level;
        const unsigned long long local = task - plan.cls[ci].task_first;
        const uint32_t rho = (uint32_t)(local / plan.cls[ci].n_blocks);
        const uint32_t blk = plan.cls[ci].blk_first + (uint32_t)(local % plan.cls[ci].n_blocks);
        uint32_t total = 0, mn = JOIN_INVALID;
#pragma unroll
        for (int w = 0; w < RW; w++) {
            const uint32_t r = w * 32u + lane;
            const uint32_t pos = plan.rows[(size_t)blk * ROWS + r];
            uint32_t id = JOIN_INVALID;
            if (pos != JOIN_INVALID) id = plan.row_ids ? plan.row_ids[pos] : pos;
            ti.rowpos[r] = pos;
            ti.rowid[r] = id;
            RangeRun<KT> rr{};
            if (id != JOIN_INVALID) rr = range_run<KT>(sets[id].main, L, rho);
            ti.run[r] = rr;
            total += rr.hi - rr.lo;
            mn = min(mn, id);
        }
        total = __reduce_add_sync(FULL, total);
        mn = __reduce_min_sync(FULL, mn);
        if (lane == 0) {
            ti.level = L;
            ti.rho = rho;
            ti.fill = total;  // checked before anything is inserted, so the table can never fill up
            ti.min_id = mn;
        }
    };
    if (warp == 0) prepare(s_ti[0]);

    for (uint32_t p = 0;; p ^= 1u) {
        __syncthreads();  // s_ti[p] is ready, and every warp is done probing the previous table
        const TaskInfo &ti = s_ti[p];
        if (ti.task >= plan.n_tasks) return;
        const uint32_t L = ti.level, rho = ti.rho;
        const uint32_t fs = (uint32_t)plan.key_bits - L;  // key bits below the range index (32-bit low words: 1..31)
        // 32-bit low words: no key of the range has this low word (bit 31 lies inside the range index and is
        // flipped); 64-bit keys: the all-ones word is never a key (KEY_SENTINEL)
        const KT EMPTY = KS == 4 ? (KT)(((uint32_t)((unsigned long long)rho << fs)) ^ 0x80000000u) : (KT)~(KT)0;
        const bool overfull = ti.fill > MAXFILL;
        if (overfull && tid == 0) atomicExch(err, 1u);  // the host redoes the call with the merge kernel
        {  // clear the table
            uint4 *k4 = (uint4 *)keys, *m4 = (uint4 *)masks;
            const uint32_t e_lo = (uint32_t)EMPTY, e_hi = KS == 4 ? (uint32_t)EMPTY : 0xFFFFFFFFu;
            const uint4 e4 = make_uint4(e_lo, e_hi, e_lo, e_hi), z4 = make_uint4(0, 0, 0, 0);
            for (uint32_t i = tid; i < SLOTS * KS / 16; i += THREADS) k4[i] = e4;
            for (uint32_t i = tid; i < RW * SLOTS / 4; i += THREADS) m4[i] = z4;
        }
        if (tid == 0) s_colgrp = 0;
        __syncthreads();

        // ---- build: the rows' runs of this range go into the table -------------------------------------
        // a warp per row; the keys of a row are requested KB per lane at a time before any is inserted
        constexpr int KB = 192 / ROWS;  // a row's run is ~fill / ROWS keys: one batch
        for (uint32_t r = warp; r < (uint32_t)ROWS && !overfull; r += NW) {
            const RangeRun<KT> rr = ti.run[r];
            const uint32_t bit = 1u << (r & 31u);
            uint32_t *mw = masks + (r >> 5) * SLOTS;
            for (uint32_t base = rr.lo; base < rr.hi; base += 32u * KB) {
                KT kk[KB];
#pragma unroll
                for (int q = 0; q < KB; q++) {
                    const uint32_t i = base + q * 32u + lane;
                    kk[q] = i < rr.hi ? ldg_key(rr.lows + i) : (KT)0;
                }
#pragma unroll
                for (int q = 0; q < KB; q++) {
                    const uint32_t i = base + q * 32u + lane;
                    if (i >= rr.hi || ((kk[q] >> fs) & rr.fm) != rr.fv) continue;
                    uint32_t s = join_off4<SLOTS_LOG2>(kk[q]) >> 2;
                    for (;;) {
                        const KT old = cas_key(&keys[s], EMPTY, kk[q]);
                        if (old == EMPTY || old == kk[q]) {
                            atomicOr(&mw[s], bit);
                            break;
                        }
                        s = (s + 1u) & SMASK;
                    }
                }
            }
        }
        __syncthreads();
        if (warp == 0) prepare(s_ti[p ^ 1u]);  // the next task's look-ups hide behind the other warps' probing
        if (overfull) continue;

        // ---- probe: stream every column's run of the range ---------------------------------------------
        // The columns are dealt in groups: group g holds the columns cbeg + g, cbeg + g + G, ... (related genomes
        // have neighbouring ids, so the expensive columns spread over all groups); a warp takes the next group,
        // prepares its columns one per lane, and requests the keys of the next column before it probes the
        // current one.
        const uint32_t cbeg = plan.mode == PAIRS_UPPER ? ti.min_id + 1u : 0u;  // upper triangle: columns right of the first row
        const uint32_t cend = plan.n_cols;
        const uint32_t ncols = cend > cbeg ? cend - cbeg : 0u;
        // about three groups per warp, at most 32 columns each
        uint32_t G = (ncols + 31u) / 32u;
        if (G < 3u * NW) G = 3u * NW;
        if (G > ncols) G = ncols;
        // lane r answers for row 32w + r in the validity ballots; the tally reduction (flush, below) leaves lane l with
        // the counts of the rows frow(l, 0..RW-1), so those are the ids it needs for the result slots
        uint32_t my_rowid[RW], f_rowid[RW], f_rowpos[RW];
#pragma unroll
        for (int w = 0; w < RW; w++) {
            my_rowid[w] = ti.rowid[w * 32 + lane];
            uint32_t fr;
            if (RW == 2) fr = 32u * ((lane >> 4) & 1u) + (((lane >> 3) & 1u) << 2 | ((lane >> 2) & 1u) << 1 | ((lane >> 1) & 1u)) + 8u * (2u * (lane & 1u) + (uint32_t)w);
            else fr = (((lane >> 4) & 1u) << 2 | ((lane >> 3) & 1u) << 1 | ((lane >> 2) & 1u)) + 8u * (2u * ((lane >> 1) & 1u) + (lane & 1u));
            f_rowid[w] = ti.rowid[fr];
            f_rowpos[w] = ti.rowpos[fr];
        }
        for (;;) {
            uint32_t g = 0;
            if (lane == 0) g = atomicAdd(&s_colgrp, 1u);
            g = __shfl_sync(FULL, g, 0);
            if (g >= G) break;
            const uint32_t c_mine = cbeg + g + lane * G;
            uint32_t col_id = 0;
            RangeRun<KT> cr{};
            if (c_mine < cend) {
                col_id = plan.col_ids ? plan.col_ids[c_mine] : c_mine;
                cr = range_run<KT>(sets[col_id].main, L, rho);
            }
            const uint32_t nb = (ncols - g + G - 1u) / G;  // columns of this group (<= 32)
            KT nk[KPT];
            uint32_t nlo, nhi;
            const KT *nlows;
            auto fetch = [&](uint32_t j) {
                nlo = __shfl_sync(FULL, cr.lo, j);
                nhi = __shfl_sync(FULL, cr.hi, j);
                nlows = (const KT *)__shfl_sync(FULL, (unsigned long long)cr.lows, j);
#pragma unroll
                for (int q = 0; q < KPT; q++) {
                    const uint32_t i = nlo + q * 32u + lane;
                    nk[q] = i < nhi ? ldg_key(nlows + i) : (KT)0;
                }
            };
            fetch(0);
            // matches: acc[w][x] holds four 8-bit counters, byte b = row 32w + x + 8b, of THIS lane's keys;
            // all zero whenever nothing is pending (flush clears them), so columns without a match never touch them
            uint32_t acc[RW][8];
#pragma unroll
            for (int w = 0; w < RW; w++)
#pragma unroll
                for (int x = 0; x < 8; x++) acc[w][x] = 0;
            for (uint32_t j = 0; j < nb; j++) {
                const uint32_t lo = nlo, hi = nhi;
                const KT *lows = nlows;
                KT k[KPT];
#pragma unroll
                for (int q = 0; q < KPT; q++) k[q] = nk[q];
                if (j + 1 < nb) fetch(j + 1);
                if (lo >= hi) continue;
                const KT fm = __shfl_sync(FULL, cr.fm, j), fv = __shfl_sync(FULL, cr.fv, j);
                const uint32_t cid = __shfl_sync(FULL, col_id, j);
                // rows that form a requested pair with this column (upper triangle: row id < column id)
                uint32_t rowmask[RW];
#pragma unroll
                for (int w = 0; w < RW; w++)
                    rowmask[w] = plan.mode == PAIRS_UPPER ? __ballot_sync(FULL, my_rowid[w] < cid) : FULL;
                uint32_t pending = 0;  // trips added to acc since the last flush (warp-uniform)
                uint32_t cnt[RW];      // lane l: matches of its rows frow(l, .) with this column in this range
#pragma unroll
                for (int w = 0; w < RW; w++) cnt[w] = 0;
                // Sum of every byte counter over the 32 lanes, as a reduce-scatter on the packed words: each step a lane
                // keeps half of its words (later: half of its bytes) and receives the partner's copy of that half, so
                // after five steps lane l holds the totals of its RW rows frow(l, .).  Per-lane counters stay <= 7
                // (PEND_MAX trips), so no byte exceeds 32 * 7 on the way.
                auto flush = [&]() {
                    constexpr int NREG = 8 * RW;
                    uint32_t r[NREG];
#pragma unroll
                    for (int w = 0; w < RW; w++)
#pragma unroll
                        for (int x = 0; x < 8; x++) {
                            r[w * 8 + x] = acc[w][x];
                            acc[w][x] = 0;
                        }
                    uint32_t bit = 16u;
#pragma unroll
                    for (int n = NREG; n > 1; n >>= 1, bit >>= 1) {
                        const bool up = (lane & bit) != 0u;
#pragma unroll
                        for (int y = 0; y < n / 2; y++) {
                            const uint32_t send = up ? r[y] : r[y + n / 2];
                            const uint32_t keep = up ? r[y + n / 2] : r[y];
                            r[y] = keep + __shfl_xor_sync(FULL, send, bit);
                        }
                    }
                    uint32_t v = r[0];
                    if (RW == 1) {  // bit == 2: keep two of the four bytes
                        const bool up = (lane & 2u) != 0u;
                        const uint32_t send = up ? (v & 0xFFFFu) : (v >> 16), keep = up ? (v >> 16) : (v & 0xFFFFu);
                        v = keep + __shfl_xor_sync(FULL, send, 2);
                        const bool up1 = (lane & 1u) != 0u;
                        const uint32_t send1 = up1 ? (v & 0xFFu) : ((v >> 8) & 0xFFu), keep1 = up1 ? ((v >> 8) & 0xFFu) : (v & 0xFFu);
                        cnt[0] += keep1 + __shfl_xor_sync(FULL, send1, 1);
                    } else {  // bit == 1: keep two of the four bytes, one per mask word position
                        const bool up = (lane & 1u) != 0u;
                        const uint32_t send = up ? (v & 0xFFFFu) : (v >> 16), keep = up ? (v >> 16) : (v & 0xFFFFu);
                        v = keep + __shfl_xor_sync(FULL, send, 1);
                        cnt[0] += v & 0xFFu;
                        cnt[RW - 1] += (v >> 8) & 0xFFu;
                    }
                    pending = 0;
                };
                constexpr uint32_t PEND_MAX = 7 / KPT > 0 ? 7 / KPT : 1;
                static_assert(KPT <= 7, "byte counters: a lane adds at most 7 per flush");
                for (uint32_t base = lo; base < hi; base += 32u * KPT) {
                    uint32_t m[RW][KPT], anym = 0;
#pragma unroll
                    for (int q = 0; q < KPT; q++) {
                        const uint32_t i = base + q * 32u + lane;
                        if (base != lo) k[q] = i < hi ? ldg_key(lows + i) : (KT)0;
                        const bool ok = i < hi && ((k[q] >> fs) & fm) == fv;
                        // every lane loads (lanes without a key probe slot hash(0), a broadcast) so the probe is
                        // branch-free; an empty slot carries mask 0, so a stray hit on it counts nothing
                        uint32_t off = join_off4<SLOTS_LOG2>(k[q]);
                        KT e = lds_key(kbase, off, k[q]);
                        e = ok ? e : EMPTY;
                        while (e != k[q] && e != EMPTY) {
                            off = (off + 4u) & M4;
                            e = lds_key(kbase, off, k[q]);
                        }
                        const bool hit = ok && e == k[q];
#pragma unroll
                        for (int w = 0; w < RW; w++) {
                            const uint32_t mv = lds_u32(kbase + off + MASKS_OFF + w * (SLOTS * 4u));
                            m[w][q] = hit ? (mv & rowmask[w]) : 0u;
                            anym |= m[w][q];
                        }
                    }
                    if (__any_sync(FULL, anym != 0u)) {
#pragma unroll
                        for (int w = 0; w < RW; w++)
#pragma unroll
                            for (int q = 0; q < KPT; q++)
#pragma unroll
                                for (int x = 0; x < 8; x++) acc[w][x] += (m[w][q] >> x) & 0x01010101u;
                        pending++;
                    }
                    // one call site: when the counters are due, or behind the column's last trip
                    if (pending && (pending == PEND_MAX || base + 32u * KPT >= hi)) flush();
                }
#pragma unroll
                for (int w = 0; w < RW; w++) {
                    if (!cnt[w]) continue;
                    unsigned long long t;
                    bool valid = true;
                    if (plan.mode == PAIRS_UPPER) {
                        const unsigned long long i = f_rowid[w], n = plan.n_cols;
                        t = i * (2ull * n - i - 1ull) / 2ull + ((unsigned long long)cid - i - 1ull) - plan.first;
                        valid = t < plan.count;  // pairs before `first` wrap around to huge values
                    } else {
                        t = (unsigned long long)f_rowpos[w] * plan.stride_r +
                            (unsigned long long)(cbeg + g + j * G) * plan.stride_c;
                    }
                    if (valid) atomicAdd(&counts[t], cnt[w]);
                }
            }
        }
    }
}

// Geometries: <log2 slots, threads per CTA, CTAs per SM, rows per block, keys per lane per trip>; a slot is a key
// word plus one 4-byte mask word per 32 rows.  The first N_JCFG32 entries serve 32-bit low words, the rest 64-bit keys.
#define GKD_FOR_EACH_JCFG32(X)                                                                                   \
    X(0, 14, 1024, 1, 64, 3) X(1, 14, 1024, 1, 32, 6) X(2, 13, 512, 2, 32, 3) X(3, 14, 512, 1, 64, 3) X(4, 13, 512, 2, 64, 2) \
    X(5, 12, 256, 4, 32, 2) X(6, 14, 1024, 1, 64, 4) X(7, 14, 768, 1, 64, 3)
#define GKD_FOR_EACH_JCFG64(X) X(8, 13, 768, 1, 64, 2) X(9, 13, 512, 2, 32, 3) X(10, 13, 1024, 1, 64, 2) X(11, 12, 512, 2, 64, 1)
constexpr int N_JCFG32 = 8, N_JCFG = 12;
static const int g_jcfg_slots_log2[N_JCFG] = {14, 14, 13, 14, 13, 12, 14, 14, 13, 13, 13, 12};
static const int g_jcfg_rows[N_JCFG] = {64, 32, 32, 64, 64, 32, 64, 64, 64, 32, 64, 64};

// GKD_JOIN_CFG pins a geometry (of the right key width); otherwise 64-row blocks when the call has enough rows and
// columns to fill them (measured on the B200, tools/bench_rect.py: 125 x 375 and 300 x 300 blocks are 4-7 % faster
// with 64 rows), 32-row blocks for smaller calls
int join_pick_cfg(uint32_t n_rows, uint32_t n_cols, int low_bits) {
    const int first = low_bits == 32 ? 0 : N_JCFG32, last = low_bits == 32 ? N_JCFG32 : N_JCFG;
    if (const char *e = getenv("GKD_JOIN_CFG")) {
        const int c = atoi(e);
        if (c >= first && c < last) return c;
    }
    const bool big = n_rows >= 96 && n_cols >= 128;
    return low_bits == 32 ? (big ? 7 : 1) : (big ? 8 : 9);
}
uint32_t join_cfg_slots(int cfg) { return 1u << g_jcfg_slots_log2[cfg]; }
uint32_t join_cfg_rows(int cfg) { return (uint32_t)g_jcfg_rows[cfg]; }

cudaError_t join_configure() {
    cudaError_t e;
#define X(i, SL, T, C, R, K)                                                                                   \
    if ((e = cudaFuncSetAttribute(k_join<uint32_t, SL, T, C, R, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (4 << SL) * (1 + R / 32))) != cudaSuccess)                                   \
        return e;
    GKD_FOR_EACH_JCFG32(X)
#undef X
#define X(i, SL, T, C, R, K)                                                                                   \
    if ((e = cudaFuncSetAttribute(k_join<uint64_t, SL, T, C, R, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (4 << SL) * (2 + R / 32))) != cudaSuccess)                                   \
        return e;
    GKD_FOR_EACH_JCFG64(X)
#undef X
    return cudaSuccess;
}

cudaError_t launch_join(const SetDesc *sets, const JoinPlan &plan, uint32_t *counts, unsigned long long *work_counter,
                        uint32_t *err, int n_sms, cudaStream_t s) {
    if (plan.n_tasks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    const int c = plan.cfg;
#define LAUNCH(KT, KW, i, SL, T, C, R, K)                                                                       \
    if (c == i) {                                                                                               \
        unsigned long long grid = (unsigned long long)n_sms * C;                                                \
        if (grid > plan.n_tasks) grid = plan.n_tasks;                                                           \
        k_join<KT, SL, T, C, R, K><<<(unsigned)grid, T, (4 << SL) * (KW + R / 32), s>>>(sets, plan, counts, work_counter, err); \
        return cudaGetLastError();                                                                              \
    }
#define X(i, SL, T, C, R, K) LAUNCH(uint32_t, 1, i, SL, T, C, R, K)
    GKD_FOR_EACH_JCFG32(X)
#undef X
#define X(i, SL, T, C, R, K) LAUNCH(uint64_t, 2, i, SL, T, C, R, K)
    GKD_FOR_EACH_JCFG64(X)
#undef X
#undef LAUNCH
    return cudaErrorInvalidValue;
}

}  // namespace gkd
