// sort_unique.cu -- kernel 3 of libgkd.so: hand-written segmented LSD radix sort of every genome's
// k-mer keys followed by a unique/compaction pass that emits the sorted canonical key set plus the
// sorted list of its reverse-palindromic members.  sm_100a only; no CUB/Thrust.
//
// Reference semantics restated: HashSet<String>.add ignores duplicates, so a genome's set is the
// DISTINCT k-mers (KmerCountProcessor.java:76-77 iterates the set; SequenceKmers.distance works on
// set sizes, FastaDistanceProcessor.java:186).  Sorting + adjacent-unique gives the same set exactly.
//
// Bound: HBM.  Ideal algorithmic bytes are 16 B/key (read once, write once); this LSD implementation
// physically moves 24 B/key per pass (8 B histogram read + 8 B read + 8 B write) over
// ceil(key_bits / 8) passes with evened-out digit widths (42-bit keys: six 7-bit passes, which keeps
// the per-digit write runs of a 4096-key tile at 256 B; 9-bit digits / 5 passes measured no faster
// because the runs shrink to 64 B), plus 8 B (count) + 16 B (compact write) for the unique pass.
#include "gkd_internal.cuh"

namespace gkd {

constexpr int SORT_WARPS = SORT_THREADS / 32;

// ---- per-tile digit histogram -------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
    k_radix_hist(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, const uint64_t *__restrict__ keys,
                 uint32_t *__restrict__ tile_hist, int shift, uint32_t dmask) {
    __shared__ uint32_t s_hist[SORT_WARPS][RADIX_BINS];
    __shared__ uint32_t s_g;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX_BINS; i += SORT_THREADS) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const uint32_t slot0 = (blockIdx.x - G.tile_first) * SORT_TILE;
    const uint32_t count = min((uint32_t)SORT_TILE, G.n_slots - slot0);
    const uint64_t *src = keys + G.raw_off + slot0;
    const int warp = threadIdx.x >> 5;
#pragma unroll 4
    for (uint32_t idx = threadIdx.x; idx < count; idx += SORT_THREADS) {
        uint32_t d = (uint32_t)(src[idx] >> shift) & dmask;
        atomicAdd(&s_hist[warp][d], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX_BINS; d += SORT_THREADS) {
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) c += s_hist[w][d];
        tile_hist[(size_t)blockIdx.x * RADIX_BINS + d] = c;
    }
}

// ---- block-wide exclusive scan of one value per thread (256 threads) -----------------------------
template <typename T, int NW = SORT_WARPS>
__device__ __forceinline__ T block_exclusive_scan(T v, T *s_warp /* [NW] */, T &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    T base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        T c = s_warp[w];
        if (w < warp) base += c;
        tot += c;
    }
    total = tot;
    __syncthreads();
    return base + incl - v;
}

// ---- per-genome scan of the tile histograms: (digit-major, tile-minor) exclusive offsets ----------
__global__ void __launch_bounds__(RADIX_BINS)
    k_radix_scan(const BatchGenome *__restrict__ genomes, uint32_t *__restrict__ tile_hist) {
    __shared__ uint32_t s_warp[RADIX_BINS / 32];
    const BatchGenome G = genomes[blockIdx.x];
    if (G.n_tiles == 0) return;
    const int d = threadIdx.x;
    uint32_t *col = tile_hist + (size_t)G.tile_first * RADIX_BINS + d;
    uint32_t run = 0;
#pragma unroll 8
    for (uint32_t t = 0; t < G.n_tiles; t++) {
        uint32_t c = col[(size_t)t * RADIX_BINS];
        col[(size_t)t * RADIX_BINS] = run;
        run += c;
    }
    uint32_t total;
    uint32_t base = block_exclusive_scan<uint32_t, RADIX_BINS / 32>(run, s_warp, total);
#pragma unroll 8
    for (uint32_t t = 0; t < G.n_tiles; t++) col[(size_t)t * RADIX_BINS] += base;
}

// ---- stable scatter of one tile by one digit -------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS, 4)
    k_radix_scatter(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, const uint64_t *__restrict__ in,
                    uint64_t *__restrict__ out, const uint32_t *__restrict__ tile_offs, int shift, uint32_t dmask) {
    extern __shared__ __align__(16) uint64_t s_keys[];  // SORT_TILE keys (dynamic: static smem is at its 48 KiB cap)
    __shared__ uint32_t s_cnt[SORT_WARPS][RADIX_BINS];
    __shared__ uint32_t s_dstart[RADIX_BINS];
    __shared__ uint32_t s_goff[RADIX_BINS];
    __shared__ uint32_t s_warp[SORT_WARPS];
    __shared__ uint32_t s_g;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX_BINS; i += SORT_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const uint32_t slot0 = (blockIdx.x - G.tile_first) * SORT_TILE;
    const uint32_t count = min((uint32_t)SORT_TILE, G.n_slots - slot0);
    const uint64_t *src = in + G.raw_off + slot0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // warp w owns tile keys [w*512, (w+1)*512); item i of lane l is key w*512 + i*32 + l, so the
    // (item, lane) order is the input order and the ranking below is stable.
    uint64_t key[SORT_ITEMS];
    uint16_t rank[SORT_ITEMS];  // < 4096, packed to keep the kernel at 64 registers (4 CTAs per SM)
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        uint32_t idx = warp * (32 * SORT_ITEMS) + i * 32 + lane;
        key[i] = idx < count ? src[idx] : KEY_SENTINEL;  // padding ranks after every real key
    }
    // Rank = (keys of the same digit seen by this warp in earlier items) + (same-digit lanes below me).
    // The first part comes from ONE shared-memory atomic per digit group per item, issued by the
    // group's lowest lane; atomics to one address retire in issue order, so no __syncwarp is needed
    // and the 16 match/atomic/shuffle chains overlap instead of serialising on shared memory.
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t d = (uint32_t)(key[i] >> shift) & dmask;
        // lanes holding the same digit: eight ballots (one per digit bit) instead of MATCH.ANY, whose
        // latency dominated this kernel (44 % of the stall samples sat on its consumers)
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < RADIX_BITS; b++) {
            const uint32_t vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? vote : ~vote;
        }
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (lane == leader) prev = atomicAdd(&s_cnt[warp][d], (uint32_t)__popc(peers));
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[i] = (uint16_t)(prev + __popc(peers & lt_mask));
    }
    __syncthreads();
    {   // per digit: exclusive scan over warps, then over digits
        constexpr int DPT = (RADIX_BINS + SORT_THREADS - 1) / SORT_THREADS;  // digits per thread
        static_assert(DPT == 1 || DPT == 2, "digit scan handles up to 2 digits per thread");
        uint32_t sum[2] = {0, 0};
#pragma unroll
        for (int h = 0; h < DPT; h++) {
            const int d = threadIdx.x + h * SORT_THREADS;
            uint32_t acc = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) {
                uint32_t c = s_cnt[w][d];
                s_cnt[w][d] = acc;
                acc += c;
            }
            sum[h] = acc;
        }
        uint32_t total0, total1;
        const uint32_t start0 = block_exclusive_scan<uint32_t>(sum[0], s_warp, total0);
        s_dstart[threadIdx.x] = start0;
        s_goff[threadIdx.x] = tile_offs[(size_t)blockIdx.x * RADIX_BINS + threadIdx.x];
        if (DPT == 2) {
            const uint32_t start1 = block_exclusive_scan<uint32_t>(sum[1], s_warp, total1) + total0;
            s_dstart[threadIdx.x + SORT_THREADS] = start1;
            s_goff[threadIdx.x + SORT_THREADS] = tile_offs[(size_t)blockIdx.x * RADIX_BINS + threadIdx.x + SORT_THREADS];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        uint32_t d = (uint32_t)(key[i] >> shift) & dmask;
        s_keys[s_dstart[d] + s_cnt[warp][d] + rank[i]] = key[i];
    }
    __syncthreads();
    uint64_t *dst = out + G.raw_off;
#pragma unroll 4
    for (uint32_t idx = threadIdx.x; idx < count; idx += SORT_THREADS) {
        uint64_t kx = s_keys[idx];
        uint32_t d = (uint32_t)(kx >> shift) & dmask;
        dst[s_goff[d] + (idx - s_dstart[d])] = kx;
    }
}

cudaError_t sort_configure() {
    // per-device attribute: every context sets it after cudaSetDevice (a process-wide flag would leave a
    // second device unconfigured)
    return cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(SORT_TILE * sizeof(uint64_t)));
}

cudaError_t launch_sort(const BatchGenome *genomes, const SortPlan &plan, uint64_t **sorted_out, uint32_t *passes,
                        cudaStream_t s) {
    uint64_t *src = plan.keys_a, *dst = plan.keys_b;
    uint32_t np = 0;
    if (plan.n_tiles > 0 && plan.key_bits > 0) {
        // fewest passes of at most RADIX_BITS bits, widths as even as possible (42 bits -> six 7-bit passes)
        const int n_pass = (plan.key_bits + RADIX_BITS - 1) / RADIX_BITS;
        const int base = plan.key_bits / n_pass, extra = plan.key_bits % n_pass;
        int shift = 0;
        for (int p = 0; p < n_pass; p++) {
            const int bits = base + (p >= n_pass - extra ? 1 : 0);
            const uint32_t dmask = (1u << bits) - 1u;
            k_radix_hist<<<plan.n_tiles, SORT_THREADS, 0, s>>>(genomes, plan.n_genomes, src, plan.tile_hist, shift, dmask);
            k_radix_scan<<<plan.n_genomes, RADIX_BINS, 0, s>>>(genomes, plan.tile_hist);
            k_radix_scatter<<<plan.n_tiles, SORT_THREADS, SORT_TILE * sizeof(uint64_t), s>>>(
                genomes, plan.n_genomes, src, dst, plan.tile_hist, shift, dmask);
            uint64_t *t = src;
            src = dst;
            dst = t;
            shift += bits;
            np++;
        }
    }
    *sorted_out = src;
    *passes = np;
    return cudaGetLastError();
}

// ---- unique / compaction into bucketed sets ----------------------------------------------------------
// packed flags of one sorted position: low word = "first occurrence of a real key", high word = "and
// the k-mer is its own reverse complement".  The sorted values are h = mix(key); the palindrome test
// needs the key itself, so it un-mixes (even K only).
__device__ __forceinline__ uint64_t unique_flags(const uint64_t *__restrict__ sorted, uint32_t idx, uint32_t n,
                                                 bool check_pal, int k, const MixParams &mix, uint64_t &h_out) {
    if (idx >= n) {
        h_out = KEY_SENTINEL;
        return 0;
    }
    uint64_t h = sorted[idx];
    h_out = h;
    if (h == KEY_SENTINEL) return 0;
    if (idx > 0 && sorted[idx - 1] == h) return 0;
    uint64_t f = 1;
    if (check_pal) {
        const uint64_t kx = unmix_key(h, mix);
        const uint64_t kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1);
        const uint64_t rc = (~(reverse_pairs(kx) >> (64 - 2 * k))) & kmask;
        if (rc == kx) f |= 1ull << 32;
    }
    return f;
}

__global__ void __launch_bounds__(SORT_THREADS)
    k_unique_count(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, const uint64_t *__restrict__ sorted,
                   uint64_t *__restrict__ tile_uniq, int check_pal, int k, MixParams mix) {
    __shared__ uint64_t s_warp[SORT_WARPS];
    __shared__ uint32_t s_g;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const uint32_t slot0 = (blockIdx.x - G.tile_first) * SORT_TILE;
    const uint64_t *src = sorted + G.raw_off;
    uint64_t acc = 0, h;
#pragma unroll 4
    for (int i = 0; i < SORT_ITEMS; i++)
        acc += unique_flags(src, slot0 + i * SORT_THREADS + threadIdx.x, G.n_slots, check_pal != 0, k, mix, h);
    uint64_t total;
    block_exclusive_scan<uint64_t>(acc, s_warp, total);
    if (threadIdx.x == 0) tile_uniq[blockIdx.x] = total;
}

// one CTA per genome: exclusive scan of its tile counts in place, totals to genome_counts
__global__ void __launch_bounds__(SORT_THREADS)
    k_unique_scan(const BatchGenome *__restrict__ genomes, uint64_t *__restrict__ tile_uniq,
                  uint64_t *__restrict__ genome_counts) {
    __shared__ uint64_t s_warp[SORT_WARPS];
    const BatchGenome G = genomes[blockIdx.x];
    uint64_t carry = 0;
    uint64_t *col = tile_uniq + G.tile_first;
    for (uint32_t t0 = 0; t0 < G.n_tiles; t0 += SORT_THREADS) {
        uint32_t t = t0 + threadIdx.x;
        uint64_t v = t < G.n_tiles ? col[t] : 0;
        uint64_t total;
        uint64_t ex = block_exclusive_scan<uint64_t>(v, s_warp, total);
        if (t < G.n_tiles) col[t] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) genome_counts[blockIdx.x] = carry;
}

// bucket of h at a table level (level 0 = one bucket)
__device__ __forceinline__ uint32_t bucket_of(uint64_t h, int key_bits, uint32_t level) {
    return level == 0 ? 0u : (uint32_t)(h >> (key_bits - (int)level));
}

// Writes the low words of the distinct keys in order and, from the same pass, the bucket offset table:
// the thread that holds the first occurrence of a key also knows the previous distinct key (its left
// neighbour in the sorted slots), so it fills the table entries of every bucket boundary between the
// two; the thread of the last distinct key fills the tail.  Palindromic members go to a scratch list of
// full h values that k_set_finish turns into the set's palindrome sub-set.
template <typename LowT>
__global__ void __launch_bounds__(SORT_THREADS)
    k_unique_write(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, const uint64_t *__restrict__ sorted,
                   const uint64_t *__restrict__ tile_uniq, const SetBuild *__restrict__ dst, int check_pal, int k,
                   MixParams mix) {
    __shared__ uint64_t s_warp[SORT_WARPS];
    __shared__ uint32_t s_g;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const SetBuild D = dst[s_g];
    LowT *lows = (LowT *)D.lows;
    const uint32_t slot0 = (blockIdx.x - G.tile_first) * SORT_TILE;
    const uint64_t *src = sorted + G.raw_off;
    uint64_t base = tile_uniq[blockIdx.x];
    const uint32_t n_buckets = 1u << D.level;
    // chunks of 256 consecutive positions keep the output order == sorted order
    for (int i = 0; i < SORT_ITEMS; i++) {
        uint64_t h;
        const uint32_t idx = slot0 + i * SORT_THREADS + threadIdx.x;
        uint64_t f = unique_flags(src, idx, G.n_slots, check_pal != 0, k, mix, h);
        uint64_t total;
        uint64_t ex = block_exclusive_scan<uint64_t>(f, s_warp, total);
        if (f & 1ull) {
            const uint64_t at64 = base + ex;
            const uint32_t at = (uint32_t)at64;
            lows[at] = (LowT)h;
            if (f >> 32) D.pal_h[(uint32_t)(at64 >> 32)] = h;
            const uint32_t b = bucket_of(h, mix.bits, D.level);
            // entries (bucket of the previous distinct key, b] all start at this key
            uint32_t q = idx == 0 ? 0u : bucket_of(src[idx - 1], mix.bits, D.level) + 1u;
            for (; q <= b; q++) D.offs[q] = at;
            if (at == D.n - 1u)
                for (q = b + 1u; q <= n_buckets; q++) D.offs[q] = D.n;
        }
        base += total;
    }
}

// one CTA per genome: the table of an empty set, and the palindrome sub-set (tiny: a k-mer is its own
// reverse complement with probability 4^-(K/2)) built from the scratch list of full h values
template <typename LowT>
__global__ void __launch_bounds__(256) k_set_finish(const SetBuild *__restrict__ dst, MixParams mix) {
    const SetBuild D = dst[blockIdx.x];
    if (D.n == 0)
        for (uint32_t q = threadIdx.x; q <= (1u << D.level); q += blockDim.x) D.offs[q] = 0;
    if (D.pal_offs == nullptr) return;
    LowT *pl = (LowT *)D.pal_lows;
    for (uint32_t i = threadIdx.x; i < D.n_pal; i += blockDim.x) pl[i] = (LowT)D.pal_h[i];
    for (uint32_t q = threadIdx.x; q <= (1u << D.pal_level); q += blockDim.x) {
        // number of palindromic members whose bucket is < q
        uint32_t lo = 0, hi = D.n_pal;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (bucket_of(D.pal_h[mid], mix.bits, D.pal_level) < q) lo = mid + 1;
            else hi = mid;
        }
        D.pal_offs[q] = lo;
    }
}

cudaError_t launch_unique_count(const BatchGenome *genomes, const SortPlan &plan, const uint64_t *sorted, int alphabet,
                                int k, MixParams mix, cudaStream_t s) {
    if (plan.n_genomes == 0) return cudaSuccess;
    int check_pal = (alphabet != GKD_PROT) && (k % 2 == 0);
    if (plan.n_tiles)
        k_unique_count<<<plan.n_tiles, SORT_THREADS, 0, s>>>(genomes, plan.n_genomes, sorted, plan.tile_uniq, check_pal, k,
                                                             mix);
    k_unique_scan<<<plan.n_genomes, SORT_THREADS, 0, s>>>(genomes, plan.tile_uniq, plan.genome_counts);
    return cudaGetLastError();
}

cudaError_t launch_unique_write(const BatchGenome *genomes, const SortPlan &plan, const uint64_t *sorted,
                                const SetBuild *dst, int alphabet, int k, MixParams mix, int low_bits, cudaStream_t s) {
    if (plan.n_genomes == 0) return cudaSuccess;
    int check_pal = (alphabet != GKD_PROT) && (k % 2 == 0);
    if (low_bits == 32) {
        if (plan.n_tiles)
            k_unique_write<uint32_t><<<plan.n_tiles, SORT_THREADS, 0, s>>>(genomes, plan.n_genomes, sorted, plan.tile_uniq,
                                                                           dst, check_pal, k, mix);
        k_set_finish<uint32_t><<<plan.n_genomes, 256, 0, s>>>(dst, mix);
    } else {
        if (plan.n_tiles)
            k_unique_write<uint64_t><<<plan.n_tiles, SORT_THREADS, 0, s>>>(genomes, plan.n_genomes, sorted, plan.tile_uniq,
                                                                           dst, check_pal, k, mix);
        k_set_finish<uint64_t><<<plan.n_genomes, 256, 0, s>>>(dst, mix);
    }
    return cudaGetLastError();
}

// ---- import / export helpers -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mix_keys(uint64_t *__restrict__ keys, uint64_t n, MixParams mix) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t kx = keys[i];
        // a value outside the key space, or the all-ones key (never a valid k-mer, see MixParams::fix), cannot
        // be a member of a set of this context: drop it (the sentinel sorts last)
        keys[i] = ((kx & ~mix.mask) || kx == mix.mask) ? KEY_SENTINEL : mix_key(kx, mix);
    }
}

cudaError_t launch_mix_keys(uint64_t *keys, uint64_t n, MixParams mix, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_mix_keys<<<(unsigned)blocks, 256, 0, s>>>(keys, n, mix);
    return cudaGetLastError();
}

// one thread per key: find its bucket by binary search in the offset table, rebuild h, un-mix
template <typename LowT>
__global__ void __launch_bounds__(256) k_unmix_set(SubSet S, MixParams mix, uint64_t *__restrict__ keys_out) {
    const LowT *lows = (const LowT *)S.lows;
    const int rest = mix.bits - (int)S.level;  // bits of h below the bucket bits
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
        uint64_t h = (uint64_t)lows[i];
        if (sizeof(LowT) == 4 && S.level > 0) {
            // largest bucket b with offs[b] <= i
            uint32_t lo = 0, hi = 1u << S.level;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (S.offs[mid] <= i) lo = mid;
                else hi = mid;
            }
            h = ((uint64_t)lo << rest) | (h & ((1ull << rest) - 1ull));
        }
        keys_out[i] = unmix_key(h, mix);
    }
}

cudaError_t launch_unmix_set(SubSet set, MixParams mix, int low_bits, uint64_t *keys_out, cudaStream_t s) {
    if (set.n == 0) return cudaSuccess;
    uint32_t blocks = (set.n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (low_bits == 32) k_unmix_set<uint32_t><<<blocks, 256, 0, s>>>(set, mix, keys_out);
    else k_unmix_set<uint64_t><<<blocks, 256, 0, s>>>(set, mix, keys_out);
    return cudaGetLastError();
}

__global__ void k_fill_u64(uint64_t *dst, uint64_t n, uint64_t value) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = value;
}

cudaError_t launch_fill_u64(uint64_t *dst, uint64_t n, uint64_t value, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    k_fill_u64<<<(unsigned)blocks, 256, 0, s>>>(dst, n, value);
    return cudaGetLastError();
}

}  // namespace gkd
