// sort_msd.cu -- kernel 3, fast path: bucket sort of the mixed keys.
//
// Reference semantics restated: HashSet<String>.add ignores duplicates, so a genome's set is its DISTINCT
// k-mers (KmerCountProcessor.java:76-77; set sizes feed SequenceKmers.distance, FastaDistanceProcessor.java:186).
//
// The slots hold h = mix(key), uniform over the key space whatever the genome (gkd_internal.cuh), so an MSD
// partition by the top p bits splits a genome into bins of predictable size.  The bins therefore get FIXED
// capacity regions (mean + 8 sigma + 64), which removes the histogram and scan passes of a classic radix
// partition:
//   pass 1  k_encode_scatter (pack_encode.cu)  kernel 2 keeps its keys in registers, ranks them by bin with
//           shared-memory atomics, reserves room with one global atomic per (tile, bin) and writes every key
//           straight into its bin; invalid slots are dropped.  (k_keys_scatter does the same for imported keys.)
//   pass 2  k_msd_binsort  one CTA per bin: the bin's keys are ranked into 4096 sort-buckets (next bits of h) in
//           shared memory; every key then finds its place inside its 1-2-key sort-bucket by counting the smaller
//           ones (no per-thread insertion sort); adjacent-difference flags drop the duplicates;
//           the bin's position in the set comes from a decoupled look-back over the genome's earlier bins; the
//           low words AND the bucket offset table of the finished set go straight into the set arena.
// Traffic: 0.4 (packed codes) + 8 (keys into bins) + 8 + 4.1 (bin sort) = 20.5 B per k-mer against ~170 for
// the LSD path (8 + six passes of 24 + unique); the ideal is 16.
// Fallbacks (the LSD path of sort_unique.cu): even-K nucleotide contexts (palindrome side lists) and any batch
// in which a bin overflows its region (heavily repeated k-mers: all copies of a key share a bin).
#include "gkd_internal.cuh"

namespace gkd {

constexpr int MSD_THREADS = 512;                                            // bin-sort CTA
constexpr int MSD_KPT = (MSD_BIN_CAP + MSD_THREADS - 1) / MSD_THREADS;      // keys per thread

// per genome: keys that reached the bins, and its fullest bin (host: arena sizing, fallback decision)
__global__ void __launch_bounds__(256)
    k_msd_totals(const MsdGenome *__restrict__ msd, const uint32_t *__restrict__ bin_cursor,
                 uint32_t *__restrict__ genome_valid, uint32_t *__restrict__ genome_maxbin) {
    __shared__ uint32_t s_sum, s_max;
    const MsdGenome M = msd[blockIdx.x];
    if (threadIdx.x == 0) s_sum = 0, s_max = 0;
    __syncthreads();
    uint32_t sum = 0, mx = 0;
    for (uint32_t b = threadIdx.x; b < (1u << M.p); b += 256) {
        const uint32_t v = bin_cursor[M.bin_first + b];
        sum += v;
        mx = v > mx ? v : mx;
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_sum, sum);
        atomicMax(&s_max, mx);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        genome_valid[blockIdx.x] = s_sum;
        genome_maxbin[blockIdx.x] = s_max;
    }
}

// block-wide exclusive scan of one value per thread (MSD_THREADS threads); *total = sum
__device__ __forceinline__ uint32_t msd_block_scan(uint32_t v, uint32_t *s_warp, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < MSD_THREADS / 32; w++) {
        const uint32_t c = s_warp[w];
        if (w < warp) base += c;
        tot += c;
    }
    *total = tot;
    __syncthreads();
    return base + incl - v;
}

// look-back descriptor: bits 63..62 = state (0 empty, 1 aggregate = this bin only, 2 prefix = all bins up to
// and including this one), bits 61..0 = distinct-key count
constexpr unsigned long long LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_MASK = (1ull << 62) - 1;

// KT = what a key looks like in shared memory: its bits below the bin bits as uint32_t when they fit
// (key_bits - p <= 32 for every genome of the batch), else the whole h
template <typename KT, typename LowT>
__global__ void __launch_bounds__(MSD_THREADS, 2)
    k_msd_binsort(const MsdGenome *__restrict__ msd, uint32_t n_genomes, int key_bits, const uint32_t *__restrict__ bin_cursor,
                  const uint64_t *__restrict__ bins, unsigned long long *__restrict__ status,
                  uint32_t *__restrict__ genome_unique) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    KT *s_k = reinterpret_cast<KT *>(s_raw);   // keys grouped by sort-bucket, later the distinct keys in order
    KT *s_o = s_k + MSD_BIN_CAP;               // keys in sorted order (duplicates still there)
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_o + MSD_BIN_CAP);  // [2^s] keys per sort-bucket
    uint32_t *s_start = s_cnt + (1u << MSD_SORT_BITS);                  // [2^s + 1] start of every sort-bucket
    __shared__ uint32_t s_warp[MSD_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    // which genome owns this bin (bin_first ascending)
    uint32_t lo = 0, hi = n_genomes;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (msd[mid].bin_first <= blockIdx.x) lo = mid;
        else hi = mid;
    }
    const MsdGenome M = msd[lo];
    const uint32_t bin = blockIdx.x - M.bin_first, n_bins = 1u << M.p;
    uint32_t c = bin_cursor[blockIdx.x];
    if (c > M.cap || c > MSD_BIN_CAP) c = 0;  // never launched like this (the host falls back); keeps the look-back alive
    const uint64_t *src = bins + M.bins_off + (uint64_t)bin * M.cap;
    const int rest = key_bits - (int)M.p;                                // bits of h below the bin bits
    const uint64_t rest_mask = rest >= 64 ? ~0ull : ((1ull << rest) - 1ull);
    const uint32_t s = M.s, n_sb = 1u << s;
    const int sb_shift = rest - (int)s;                                   // >= 0

    for (uint32_t i = threadIdx.x; i < n_sb; i += MSD_THREADS) s_cnt[i] = 0;
    __syncthreads();
    // load (coalesced) + rank inside the sort-bucket (arbitrary order); keys stay in registers
    KT kv[MSD_KPT];
    uint32_t meta[MSD_KPT];  // sort-bucket | rank << 16
    const int rounds = (int)((c + MSD_THREADS - 1) / MSD_THREADS);  // the unrolled sweeps stop after the bin's last key
    uint64_t raw[MSD_KPT];
#pragma unroll
    for (int j = 0; j < MSD_KPT; j++) {  // every load of the bin in flight before the first shared-memory atomic
        if (j >= rounds) break;
        const uint32_t i = j * MSD_THREADS + threadIdx.x;
        raw[j] = i < c ? src[i] : 0ull;
    }
#pragma unroll
    for (int j = 0; j < MSD_KPT; j++) {
        if (j >= rounds) break;
        const uint32_t i = j * MSD_THREADS + threadIdx.x;
        kv[j] = 0;
        meta[j] = 0;
        if (i < c) {
            const uint64_t h = raw[j] & rest_mask;
            const uint32_t sb = s ? (uint32_t)(h >> sb_shift) : 0u;
            kv[j] = (KT)h;
            meta[j] = sb | (atomicAdd(&s_cnt[sb], 1u) << 16);
        }
    }
    __syncthreads();
    {   // exclusive scan of the sort-bucket sizes: SB_PT neighbouring values per thread
        constexpr uint32_t SB_PT = (1u << MSD_SORT_BITS) / MSD_THREADS;
        static_assert(SB_PT * MSD_THREADS == (1u << MSD_SORT_BITS) && SB_PT >= 1, "whole sort-buckets per thread");
        const uint32_t a = SB_PT * threadIdx.x;
        uint32_t v[SB_PT], sum = 0, total;
#pragma unroll
        for (uint32_t q = 0; q < SB_PT; q++) {
            v[q] = a + q < n_sb ? s_cnt[a + q] : 0u;
            sum += v[q];
        }
        uint32_t ex = msd_block_scan(sum, s_warp, &total);
#pragma unroll
        for (uint32_t q = 0; q < SB_PT; q++) {
            if (a + q < n_sb) s_start[a + q] = ex;
            ex += v[q];
        }
        if (threadIdx.x == 0) s_start[n_sb] = c;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MSD_KPT; j++) {
        if (j >= rounds) break;
        if (j * MSD_THREADS + threadIdx.x < c) s_k[s_start[meta[j] & 0xFFFFu] + (meta[j] >> 16)] = kv[j];
    }
    __syncthreads();
    // place every key inside its sort-bucket by counting: smaller keys, and equal keys that sit before it
#pragma unroll
    for (int j = 0; j < MSD_KPT; j++) {
        if (j >= rounds) break;
        if (j * MSD_THREADS + threadIdx.x < c) {
            const uint32_t sb = meta[j] & 0xFFFFu, b0 = s_start[sb], b1 = s_start[sb + 1], me = b0 + (meta[j] >> 16);
            const KT key = kv[j];
            uint32_t at = b0;
            for (uint32_t t = b0; t < b1; t++) {
                const KT v = s_k[t];
                at += (v < key || (v == key && t < me)) ? 1u : 0u;
            }
            s_o[at] = key;
        }
    }
    __syncthreads();
    // distinct keys: adjacent-difference flags over the sorted bin, contiguous chunk per thread
    const uint32_t per = (c + MSD_THREADS - 1) / MSD_THREADS;
    const uint32_t i0 = min(threadIdx.x * per, c), i1 = min(i0 + per, c);
    uint32_t mine = 0;
    for (uint32_t i = i0; i < i1; i++) mine += (i == 0 || s_o[i] != s_o[i - 1]) ? 1u : 0u;
    uint32_t u;
    uint32_t at = msd_block_scan(mine, s_warp, &u);
    for (uint32_t i = i0; i < i1; i++)
        if (i == 0 || s_o[i] != s_o[i - 1]) s_k[at++] = s_o[i];  // s_k is free again: the distinct keys, in order
    // decoupled look-back over the earlier bins of this genome (warp 0, 32 predecessors per round trip): where
    // does this bin start in the set?
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        unsigned long long prefix = 0;
        if (bin == 0) {
            if (lane == 0) atomicExch(&status[blockIdx.x], LB_PREFIX | u);
        } else {
            if (lane == 0) atomicExch(&status[blockIdx.x], LB_AGG | u);
            int top = (int)blockIdx.x - 1;  // newest predecessor not yet accounted for
            for (;;) {
                const int j = top - (int)lane;
                const bool mine_valid = j >= (int)M.bin_first;
                unsigned long long st = LB_PREFIX;  // lanes before the genome's first bin read as an empty prefix
                if (mine_valid) {
                    do {
                        st = atomicAdd(&status[j], 0ull);  // coherent read at L2
                    } while ((st >> 62) == 0);
                }
                // lanes up to and including the first PREFIX contribute
                const uint32_t is_prefix = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                const uint32_t upto = is_prefix ? (uint32_t)__ffs(is_prefix) : 32u;  // number of contributing lanes
                unsigned long long val = lane < upto ? (st & LB_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                prefix += val;
                if (is_prefix) break;
                top -= 32;
            }
            if (lane == 0) atomicExch(&status[blockIdx.x], LB_PREFIX | (prefix + u));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    const uint32_t prefix = (uint32_t)s_prefix;
    // low words of the finished set (the bin bits are put back before truncating to LowT)
    LowT *lows = (LowT *)M.lows;
    const uint64_t bin_bits = rest >= 64 ? 0ull : ((uint64_t)bin << rest);
    for (uint32_t i = threadIdx.x; i < u; i += MSD_THREADS) lows[prefix + i] = (LowT)(bin_bits | (uint64_t)s_k[i]);
    // bucket table: this bin covers the real buckets [bin << d, (bin + 1) << d), d = level - p; a real bucket starts
    // at the first distinct key whose bits below the bin reach r << (rest - d)
    const uint32_t d = M.level - M.p;
    for (uint32_t r = threadIdx.x; r < (1u << d); r += MSD_THREADS) {
        const uint64_t bound = d ? ((uint64_t)r << (rest - (int)d)) : 0ull;
        uint32_t a = 0, b = u;
        while (a < b) {
            const uint32_t mid = (a + b) >> 1;
            if ((uint64_t)s_k[mid] < bound) a = mid + 1;
            else b = mid;
        }
        M.offs[((size_t)bin << d) + r] = prefix + a;
    }
    if (bin == n_bins - 1 && threadIdx.x == 0) {
        M.offs[(size_t)1 << M.level] = prefix + u;
        genome_unique[lo] = prefix + u;
    }
}

// ---- launchers -------------------------------------------------------------------------------------------
template <typename KT>
static constexpr uint32_t binsort_smem() { return 2 * MSD_BIN_CAP * (uint32_t)sizeof(KT) + (2 * (1u << MSD_SORT_BITS) + 8) * 4; }

cudaError_t msd_configure() {
    cudaError_t e;
#define X(KT, LT)                                                                                                  \
    if ((e = cudaFuncSetAttribute(k_msd_binsort<KT, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)binsort_smem<KT>()))) \
        return e;
    X(uint32_t, uint32_t) X(uint64_t, uint32_t) X(uint64_t, uint64_t)
#undef X
    return cudaSuccess;
}

cudaError_t launch_msd_totals(const MsdGenome *msd, uint32_t n_genomes, const MsdPlan &plan, cudaStream_t s) {
    if (n_genomes == 0) return cudaSuccess;
    k_msd_totals<<<n_genomes, 256, 0, s>>>(msd, plan.bin_cursor, plan.genome_valid, plan.genome_maxbin);
    return cudaGetLastError();
}

cudaError_t launch_msd_binsort(const MsdGenome *msd, uint32_t n_genomes, const MsdPlan &plan, int low_bits, bool narrow,
                               cudaStream_t s) {
    if (plan.n_bins == 0) return cudaSuccess;
    if (low_bits == 32 && narrow)
        k_msd_binsort<uint32_t, uint32_t><<<plan.n_bins, MSD_THREADS, binsort_smem<uint32_t>(), s>>>(
            msd, n_genomes, plan.key_bits, plan.bin_cursor, plan.bins, plan.status, plan.genome_unique);
    else if (low_bits == 32)
        k_msd_binsort<uint64_t, uint32_t><<<plan.n_bins, MSD_THREADS, binsort_smem<uint64_t>(), s>>>(
            msd, n_genomes, plan.key_bits, plan.bin_cursor, plan.bins, plan.status, plan.genome_unique);
    else
        k_msd_binsort<uint64_t, uint64_t><<<plan.n_bins, MSD_THREADS, binsort_smem<uint64_t>(), s>>>(
            msd, n_genomes, plan.key_bits, plan.bin_cursor, plan.bins, plan.status, plan.genome_unique);
    return cudaGetLastError();
}

}  // namespace gkd
