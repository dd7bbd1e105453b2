// sort_msd.cu -- kernel 3, fast path: bucket sort of the mixed keys.
//
// Reference semantics restated: HashSet<String>.add ignores duplicates, so a genome's set is its DISTINCT
// k-mers (KmerCountProcessor.java:76-77; set sizes feed SequenceKmers.distance, FastaDistanceProcessor.java:186).
//
// The slots hold h = mix(key), uniform over the key space whatever the genome (gkd_internal.cuh), so an MSD
// partition by the top bits splits a genome into equal bins without looking at the data distribution:
//   pass A  k_msd_hist     per-genome histogram of the top P bits (P chosen so a bin averages <= 4096 keys)
//   pass B  k_msd_scatter  every tile ranks its keys by bin in shared memory, reserves room in each bin with
//                          one global atomic per (tile, bin) and writes them there; invalid slots are dropped;
//                          the order inside a bin is arbitrary (the next pass sorts it)
//   pass C  k_msd_binsort  one CTA per bin: load the bin, counting-scatter it by the next S bits into ~8-key
//                          sort-buckets in shared memory, one thread sorts and de-duplicates each of them, the
//                          bin's distinct keys are compacted, its position in the set comes from a decoupled
//                          look-back over the genome's earlier bins, and the low words AND the bucket offset
//                          table of the finished set are written straight into the set arena.
// Traffic: 8 (encode write) + 8 (A read) + 16 (B) + 8 + 4.1 (C) = 44 B per key against ~170 for the LSD path
// (six passes of 24 B plus the unique passes); the ideal is 16.
// Fallbacks (the LSD path of sort_unique.cu): even-K nucleotide contexts (palindrome side lists), and any
// batch in which a bin outgrows the shared-memory capacity (heavily repeated k-mers: all copies of a key
// share a bin).
#include "gkd_internal.cuh"

namespace gkd {

constexpr int MSD_THREADS = 512;  // bin-sort CTA

// ---- pass A ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
    k_msd_hist(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, const MsdGenome *__restrict__ msd,
               const uint64_t *__restrict__ keys, int key_bits, uint32_t *__restrict__ bin_count) {
    extern __shared__ uint32_t s_hist[];  // 2^p counters
    __shared__ uint32_t s_g;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const MsdGenome M = msd[s_g];
    const uint32_t n_bins = 1u << M.p;
    for (uint32_t i = threadIdx.x; i < n_bins; i += SORT_THREADS) s_hist[i] = 0;
    __syncthreads();
    const uint32_t slot0 = (blockIdx.x - G.tile_first) * SORT_TILE;
    const uint32_t count = min((uint32_t)SORT_TILE, G.n_slots - slot0);
    const uint64_t *src = keys + G.raw_off + slot0;
    const int shift = key_bits - (int)M.p;
#pragma unroll 4
    for (uint32_t idx = threadIdx.x; idx < count; idx += SORT_THREADS) {
        const uint64_t h = src[idx];
        if (h != KEY_SENTINEL) atomicAdd(&s_hist[M.p ? (uint32_t)(h >> shift) : 0u], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_bins; i += SORT_THREADS)
        if (s_hist[i]) atomicAdd(&bin_count[M.bin_first + i], s_hist[i]);
}

// one CTA per genome: exclusive scan of its bin counts -> absolute start of every bin in the scattered key
// buffer; totals and the largest bin go to the host (arena sizing, fallback decision)
__global__ void __launch_bounds__(256)
    k_msd_scan(const BatchGenome *__restrict__ genomes, const MsdGenome *__restrict__ msd,
               const uint32_t *__restrict__ bin_count, uint64_t *__restrict__ bin_start,
               uint32_t *__restrict__ genome_valid, uint32_t *__restrict__ genome_maxbin) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_carry, s_max;
    const BatchGenome G = genomes[blockIdx.x];
    const MsdGenome M = msd[blockIdx.x];
    const uint32_t n_bins = 1u << M.p;
    if (threadIdx.x == 0) s_carry = 0, s_max = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t b0 = 0; b0 < n_bins; b0 += 256) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < n_bins ? bin_count[M.bin_first + b] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        atomicMax(&s_max, v);
        __syncthreads();
        uint32_t base = s_carry;
        for (int w = 0; w < warp; w++) base += s_warp[w];
        if (b < n_bins) bin_start[M.bin_first + b] = G.raw_off + base + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        genome_valid[blockIdx.x] = s_carry;
        genome_maxbin[blockIdx.x] = s_max;
    }
}

// ---- pass B ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
    k_msd_scatter(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, const MsdGenome *__restrict__ msd,
                  const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int key_bits,
                  const uint64_t *__restrict__ bin_start, uint32_t *__restrict__ bin_cursor, uint32_t bins_cap) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(s_raw);                      // SORT_TILE keys, grouped by bin
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_raw + SORT_TILE * 8);       // [2^p] keys of this tile per bin
    uint32_t *s_pos = s_cnt + bins_cap;                                          // [2^p] start of the bin's run in s_keys
    uint32_t *s_dst = s_pos + bins_cap;                                          // [2^p] reserved offset inside the bin
    __shared__ uint32_t s_warp[SORT_THREADS / 32];
    __shared__ uint32_t s_g, s_carry;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const MsdGenome M = msd[s_g];
    const uint32_t n_bins = 1u << M.p;
    for (uint32_t i = threadIdx.x; i < n_bins; i += SORT_THREADS) s_cnt[i] = 0;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t slot0 = (blockIdx.x - G.tile_first) * SORT_TILE;
    const uint32_t count = min((uint32_t)SORT_TILE, G.n_slots - slot0);
    const uint64_t *src = in + G.raw_off + slot0;
    const int shift = key_bits - (int)M.p;
    uint64_t key[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t idx = i * SORT_THREADS + threadIdx.x;
        key[i] = idx < count ? src[idx] : KEY_SENTINEL;
        rank[i] = 0;
        if (key[i] != KEY_SENTINEL) rank[i] = atomicAdd(&s_cnt[M.p ? (uint32_t)(key[i] >> shift) : 0u], 1u);
    }
    __syncthreads();
    // exclusive scan of the per-bin counts (run starts inside the tile) + one global reservation per bin
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t b0 = 0; b0 < n_bins; b0 += SORT_THREADS) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < n_bins ? s_cnt[b] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t base = s_carry;
        for (int w = 0; w < warp; w++) base += s_warp[w];
        if (b < n_bins) {
            s_pos[b] = base + incl - v;
            s_dst[b] = v ? atomicAdd(&bin_cursor[M.bin_first + b], v) : 0u;
        }
        __syncthreads();
        if (threadIdx.x == SORT_THREADS - 1) s_carry = base + incl;
        __syncthreads();
    }
    const uint32_t n_valid = s_carry;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++)
        if (key[i] != KEY_SENTINEL) s_keys[s_pos[M.p ? (uint32_t)(key[i] >> shift) : 0u] + rank[i]] = key[i];
    __syncthreads();
    // linear sweep over the grouped keys: neighbouring threads write neighbouring slots of the same bin
    for (uint32_t idx = threadIdx.x; idx < n_valid; idx += SORT_THREADS) {
        const uint64_t h = s_keys[idx];
        const uint32_t b = M.p ? (uint32_t)(h >> shift) : 0u;
        out[bin_start[M.bin_first + b] + s_dst[b] + (idx - s_pos[b])] = h;
    }
}

// ---- pass C ------------------------------------------------------------------------------------------
// look-back descriptor: bits 63..62 = state (0 empty, 1 aggregate = this bin only, 2 prefix = all bins up to
// and including this one), bits 61..0 = distinct-key count
constexpr unsigned long long LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_MASK = (1ull << 62) - 1;

template <typename LowT>
__global__ void __launch_bounds__(MSD_THREADS)
    k_msd_binsort(const MsdGenome *__restrict__ msd, uint32_t n_genomes, int key_bits, const uint32_t *__restrict__ bin_count,
                  const uint64_t *__restrict__ bin_start, const uint64_t *__restrict__ keys,
                  unsigned long long *__restrict__ status, uint32_t *__restrict__ genome_unique, uint32_t sb_cap) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t *s_a = reinterpret_cast<uint64_t *>(s_raw);                 // MSD_BIN_CAP keys as loaded, later the compacted output
    uint64_t *s_b = s_a + MSD_BIN_CAP;                                   // MSD_BIN_CAP keys grouped by sort-bucket
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_b + MSD_BIN_CAP);   // [2^s] keys per sort-bucket, later output positions
    uint32_t *s_start = s_cnt + sb_cap;                                  // [2^s + 1] start of every sort-bucket
    __shared__ uint32_t s_warp[MSD_THREADS / 32];
    __shared__ uint32_t s_carry;
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
    uint32_t c = bin_count[blockIdx.x];
    if (c > MSD_BIN_CAP) c = 0;  // never launched like this (the host falls back to the LSD path); keeps the look-back chain alive
    const uint64_t *src = keys + bin_start[blockIdx.x];
    const uint32_t n_sb = 1u << M.s;                    // sort-buckets of this bin
    const int sb_shift = key_bits - (int)M.p - (int)M.s;  // >= 0 (host guarantees p + s <= key_bits)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (uint32_t i = threadIdx.x; i < n_sb; i += MSD_THREADS) s_cnt[i] = 0;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    // load + rank inside the sort-bucket (arbitrary order) -- two strided sweeps keep the loads coalesced
    for (uint32_t i = threadIdx.x; i < c; i += MSD_THREADS) {
        const uint64_t h = src[i];
        s_a[i] = h;
        atomicAdd(&s_cnt[M.s ? ((uint32_t)(h >> sb_shift) & (n_sb - 1u)) : 0u], 1u);
    }
    __syncthreads();
    // exclusive scan of the sort-bucket sizes
    for (uint32_t b0 = 0; b0 < n_sb; b0 += MSD_THREADS) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < n_sb ? s_cnt[b] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t base = s_carry;
        for (int w = 0; w < warp; w++) base += s_warp[w];
        if (b < n_sb) {
            s_start[b] = base + incl - v;
            s_cnt[b] = 0;  // reused as the fill cursor of the scatter below
        }
        __syncthreads();
        if (threadIdx.x == MSD_THREADS - 1) s_carry = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) s_start[n_sb] = c;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < c; i += MSD_THREADS) {
        const uint64_t h = s_a[i];
        const uint32_t sb = M.s ? ((uint32_t)(h >> sb_shift) & (n_sb - 1u)) : 0u;
        s_b[s_start[sb] + atomicAdd(&s_cnt[sb], 1u)] = h;
    }
    __syncthreads();
    // every thread sorts and de-duplicates its sort-buckets in place (insertion sort: ~8 keys each)
    for (uint32_t sb = threadIdx.x; sb < n_sb; sb += MSD_THREADS) {
        const uint32_t b0 = s_start[sb], b1 = s_start[sb + 1];
        for (uint32_t i = b0 + 1; i < b1; i++) {
            const uint64_t v = s_b[i];
            uint32_t j = i;
            while (j > b0 && s_b[j - 1] > v) {
                s_b[j] = s_b[j - 1];
                j--;
            }
            s_b[j] = v;
        }
        uint32_t w = b0;
        for (uint32_t i = b0; i < b1; i++)
            if (i == b0 || s_b[i] != s_b[i - 1]) s_b[w++] = s_b[i];
        s_cnt[sb] = w - b0;  // distinct keys of this sort-bucket, at the front of its range
    }
    __syncthreads();
    // exclusive scan of the distinct counts -> position of every sort-bucket in the bin's output (kept in s_cnt)
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_sb; b0 += MSD_THREADS) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < n_sb ? s_cnt[b] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t base = s_carry;
        for (int w = 0; w < warp; w++) base += s_warp[w];
        if (b < n_sb) {
            const uint32_t at = base + incl - v, from = s_start[b];
            for (uint32_t i = 0; i < v; i++) s_a[at + i] = s_b[from + i];  // compact (s_a is free again)
            s_cnt[b] = at;
        }
        __syncthreads();
        if (threadIdx.x == MSD_THREADS - 1) s_carry = base + incl;
        __syncthreads();
    }
    const uint32_t u = s_carry;  // distinct keys of this bin
    // decoupled look-back over the earlier bins of this genome: where does this bin start in the set?
    if (threadIdx.x == 0) {
        unsigned long long prefix = 0;
        if (bin == 0) {
            atomicExch(&status[blockIdx.x], LB_PREFIX | u);
        } else {
            atomicExch(&status[blockIdx.x], LB_AGG | u);
            uint32_t j = blockIdx.x - 1;
            for (;;) {
                unsigned long long st;
                do {
                    st = atomicAdd(&status[j], 0ull);  // coherent read at L2
                } while ((st >> 62) == 0);
                prefix += st & LB_MASK;
                if ((st >> 62) == 2 || j == M.bin_first) break;
                j--;
            }
            atomicExch(&status[blockIdx.x], LB_PREFIX | (prefix + u));
        }
        s_prefix = prefix;
    }
    __syncthreads();
    const uint32_t prefix = (uint32_t)s_prefix;
    LowT *lows = (LowT *)M.lows;
    for (uint32_t i = threadIdx.x; i < u; i += MSD_THREADS) lows[prefix + i] = (LowT)s_a[i];
    // bucket table: this bin covers the real buckets [bin << d, (bin + 1) << d), d = level - p; each real bucket is
    // a whole number of sort-buckets (s >= d), so its offset is the output position of its first sort-bucket
    const uint32_t d = M.level - M.p, sd = M.s - d;
    for (uint32_t r = threadIdx.x; r < (1u << d); r += MSD_THREADS) M.offs[((size_t)bin << d) + r] = prefix + s_cnt[r << sd];
    if (bin == n_bins - 1 && threadIdx.x == 0) {
        M.offs[(size_t)1 << M.level] = prefix + u;
        genome_unique[lo] = prefix + u;
    }
}

// ---- launchers -------------------------------------------------------------------------------------------
static uint32_t scatter_smem(uint32_t max_p) { return SORT_TILE * 8 + 3 * (1u << max_p) * 4; }
static uint32_t binsort_smem(uint32_t max_s) { return 2 * MSD_BIN_CAP * 8 + (2 * (1u << max_s) + 8) * 4; }

cudaError_t msd_configure() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_msd_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem(MSD_MAX_P)))) return e;
    if ((e = cudaFuncSetAttribute(k_msd_binsort<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)binsort_smem(MSD_MAX_S)))) return e;
    if ((e = cudaFuncSetAttribute(k_msd_binsort<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)binsort_smem(MSD_MAX_S)))) return e;
    return cudaSuccess;
}

cudaError_t launch_msd_partition(const BatchGenome *genomes, uint32_t n_genomes, uint32_t n_tiles, const MsdGenome *msd,
                                 const MsdPlan &plan, cudaStream_t s) {
    if (n_genomes == 0) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaMemsetAsync(plan.bin_count, 0, (size_t)plan.n_bins * 4, s))) return e;
    if ((e = cudaMemsetAsync(plan.bin_cursor, 0, (size_t)plan.n_bins * 4, s))) return e;
    if ((e = cudaMemsetAsync(plan.status, 0, (size_t)plan.n_bins * 8, s))) return e;
    if (n_tiles)
        k_msd_hist<<<n_tiles, SORT_THREADS, (1u << plan.max_p) * 4, s>>>(genomes, n_genomes, msd, plan.keys_in, plan.key_bits,
                                                                         plan.bin_count);
    k_msd_scan<<<n_genomes, 256, 0, s>>>(genomes, msd, plan.bin_count, plan.bin_start, plan.genome_valid, plan.genome_maxbin);
    if (n_tiles)
        k_msd_scatter<<<n_tiles, SORT_THREADS, scatter_smem(plan.max_p), s>>>(genomes, n_genomes, msd, plan.keys_in, plan.keys_out,
                                                                               plan.key_bits, plan.bin_start, plan.bin_cursor,
                                                                               1u << plan.max_p);
    return cudaGetLastError();
}

cudaError_t launch_msd_binsort(const MsdGenome *msd, uint32_t n_genomes, const MsdPlan &plan, int low_bits, cudaStream_t s) {
    if (plan.n_bins == 0) return cudaSuccess;
    const uint32_t smem = binsort_smem(plan.max_s), sb_cap = 1u << plan.max_s;
    if (low_bits == 32)
        k_msd_binsort<uint32_t><<<plan.n_bins, MSD_THREADS, smem, s>>>(msd, n_genomes, plan.key_bits, plan.bin_count, plan.bin_start,
                                                                       plan.keys_out, plan.status, plan.genome_unique, sb_cap);
    else
        k_msd_binsort<uint64_t><<<plan.n_bins, MSD_THREADS, smem, s>>>(msd, n_genomes, plan.key_bits, plan.bin_count, plan.bin_start,
                                                                       plan.keys_out, plan.status, plan.genome_unique, sb_cap);
    return cudaGetLastError();
}

}  // namespace gkd
