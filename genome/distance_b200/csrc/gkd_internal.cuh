// gkd_internal.cuh -- shared constants, layouts and launcher declarations of libgkd.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gkd.h"

namespace gkd {

// ------------------------------------------------------------------------------------------------
// Layout contracts
// ------------------------------------------------------------------------------------------------
// Key that is never a valid k-mer: the canonical form of t..t is a..a = 0, and eight 0xFF bytes are
// not text.  Used for invalid k-mer slots before the sort and for the tail padding of every set.
constexpr uint64_t KEY_SENTINEL = 0xFFFFFFFFFFFFFFFFull;

// Intersect kernel geometry.  A set in HBM is `n` sorted keys followed by sentinel keys up to
// set_padded(n): the kernel stages fixed ISECT_BLK-key blocks with TMA bulk copies and relies on the
// sentinels instead of bounds checks.  ISECT_W_MAX bounds the per-round window of every kernel
// configuration (threads x keys-per-thread), so one padding rule serves them all.
constexpr int ISECT_BLK = 512;       // keys per TMA bulk copy (4 KiB)
constexpr int ISECT_W_MAX = 4096;    // largest per-round merge window of any configuration

__host__ __device__ inline uint64_t set_padded(uint64_t n) {
    // room for the window [i, i+W+1] at i == n, rounded to whole blocks
    return ((n + ISECT_W_MAX + 2 + ISECT_BLK - 1) / ISECT_BLK) * (uint64_t)ISECT_BLK;
}

// Packed residue streams.  A genome is one stream: its contigs joined by one separator position.
// DNA/RNA: 2-bit codes a=0 c=1 g=2 t=3, 32 per uint64, position p at bits [2(p%32), +2) of word p/32,
// plus a 1-bit "invalid" mask, 32 per uint32 (non-acgt characters, separators, tail padding).
// Protein: raw bytes, plus the same mask (separators and padding only).
constexpr int PACK_POS_PER_WORD = 32;
constexpr char STREAM_SEPARATOR = 0;  // byte written between contigs in the staged text

// Sort geometry
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per tile
constexpr int RADIX_BITS = 8;               // widest digit of a pass; widths are evened out (42-bit keys: six 7-bit passes)
constexpr int RADIX_BINS = 1 << RADIX_BITS;  // row stride of the per-tile histogram table

// Encode geometry (tile == sort tile so one table of tiles serves both)
constexpr int ENC_THREADS = 256;
constexpr int ENC_PER_THREAD = 16;
constexpr int ENC_TILE = ENC_THREADS * ENC_PER_THREAD;
static_assert(ENC_TILE == SORT_TILE, "encode and sort share the tile table");

// One genome inside a build batch (device-visible).
struct BatchGenome {
    const uint64_t *codes;   // DNA: packed 2-bit words; protein: raw bytes (as uint8_t*)
    const uint32_t *mask;    // invalid bits
    uint64_t raw_off;        // offset (keys) of this genome's slots in the batch key buffers
    uint32_t n_pos;          // stream positions
    uint32_t n_slots;        // k-mer slots = max(0, n_pos - k + 1)
    uint32_t tile_first;     // first tile index of this genome in the batch tile table
    uint32_t n_tiles;
};

// Device-visible descriptor of a finished set.
struct SetDesc {
    const uint64_t *keys;
    uint32_t n;
    uint32_t n_pal;
    const uint64_t *pal_keys;  // sorted palindromic members (padded like a set), or nullptr
};

// Pair enumeration modes of the intersect kernel
enum PairMode : int { PAIRS_LIST = 0, PAIRS_UPPER = 1, PAIRS_RECT = 2 };

struct PairSource {
    int mode;
    uint32_t n;          // UPPER: number of sets; RECT: number of refs (columns)
    uint64_t first;      // UPPER: linear index of the first pair of this call
    uint64_t count;      // pairs in this call
    const uint32_t *a;   // LIST: a ids; RECT: query ids; UPPER: nullptr (ids are 0..n-1)
    const uint32_t *b;   // LIST: b ids; RECT: ref ids
};

// ------------------------------------------------------------------------------------------------
// Launchers (each returns the cudaError of the launch; all work is stream-ordered)
// ------------------------------------------------------------------------------------------------
// kernel 1
cudaError_t launch_pack_dna(const char *text, uint64_t n_pos, uint64_t *codes, uint32_t *mask, int rna,
                            cudaStream_t s);
cudaError_t launch_pack_prot(const char *text, uint64_t n_pos, uint8_t *codes, uint32_t *mask, cudaStream_t s);
// kernel 2
cudaError_t launch_encode(const BatchGenome *genomes, uint32_t n_genomes, uint32_t n_tiles, int alphabet, int k,
                          uint64_t *keys_out, cudaStream_t s);
// kernel 3: segmented LSD radix sort of each genome's slots, then unique/compact
struct SortPlan {
    uint32_t n_genomes, n_tiles;
    int key_bits;
    uint64_t *keys_a, *keys_b;      // ping-pong buffers (batch raw size)
    uint32_t *tile_hist;            // [n_tiles][RADIX_BINS] digit counts, reused as offsets
    uint64_t *tile_uniq;            // [n_tiles] packed (pal << 32 | uniq) counts -> exclusive offsets
    uint64_t *genome_counts;        // [n_genomes] packed totals
};
cudaError_t launch_sort(const BatchGenome *genomes, const SortPlan &plan, uint64_t **sorted_out, uint32_t *passes,
                        cudaStream_t s);
cudaError_t launch_unique_count(const BatchGenome *genomes, const SortPlan &plan, const uint64_t *sorted, int alphabet,
                                int k, cudaStream_t s);
struct UniqueDst {
    uint64_t *keys;      // destination of the unique keys (set arena)
    uint64_t *pal_keys;  // destination of palindromic keys (may be nullptr when n_pal == 0)
};
cudaError_t launch_unique_write(const BatchGenome *genomes, const SortPlan &plan, const uint64_t *sorted,
                                const UniqueDst *dst, int alphabet, int k, cudaStream_t s);
cudaError_t launch_fill_u64(uint64_t *dst, uint64_t n, uint64_t value, cudaStream_t s);
// kernels 4 + 5
cudaError_t intersect_configure();
cudaError_t launch_intersect(const SetDesc *sets, PairSource src, int use_pal, uint32_t seg_keys, uint32_t max_segs,
                             uint32_t *counts, unsigned long long *work_counter, int n_sms, int algo, cudaStream_t s);
cudaError_t launch_intersect_small(const SetDesc *sets, PairSource src, int use_pal, uint32_t *counts, int n_sms,
                                   cudaStream_t s);
uint32_t intersect_small_max_keys();  // sets up to this size take the warp-per-pair kernel
int intersect_select(uint64_t min_keys, uint64_t max_keys, int key_bits);  // streaming kernel for this workload: 0 = CTA merge path, 1 = warp-cooperative
int intersect_min_segment(int algo);   // smallest useful merge-path segment of that kernel
int intersect_items_per_sm(int algo);  // work items per SM that keep that kernel's workers busy
cudaError_t launch_epilogue(const SetDesc *sets, PairSource src, const uint32_t *counts, const uint32_t *pal_counts,
                            int both_strands, uint64_t *inter, double *dist, cudaStream_t s);
// synthetic data
cudaError_t launch_synth(char *dst, uint64_t len, uint64_t seed, uint32_t family, uint32_t member, double rate,
                         int protein, cudaStream_t s);
void synth_host(char *dst, uint64_t len, uint64_t seed, uint32_t family, uint32_t member, double rate, int protein);

#ifdef __CUDACC__
// which genome of a batch owns tile `tile` (tile_first is ascending)
__device__ __forceinline__ uint32_t find_genome(const BatchGenome *__restrict__ g, uint32_t n, uint32_t tile) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (g[mid].tile_first <= tile) lo = mid;
        else hi = mid;
    }
    return lo;
}

// reverse the order of the 32 two-bit groups of x (base order of a packed k-mer)
__device__ __forceinline__ uint64_t reverse_pairs(uint64_t x) {
    uint64_t y = __brevll(x);
    return ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
}
#endif

// decode the linear index of the row-major strict upper triangle of an n x n matrix
__host__ __device__ inline void upper_pair(uint64_t t, uint32_t n, uint32_t &i, uint32_t &j) {
    // row i starts at S(i) = i*(2n-i-1)/2 ; solve with a float guess and fix up
    double nn = (double)n - 0.5;
    double disc = nn * nn - 2.0 * (double)t;
    if (disc < 0) disc = 0;
    long long ii = (long long)(nn - sqrt(disc));
    if (ii < 0) ii = 0;
    if (ii > (long long)n - 2) ii = (long long)n - 2;
    auto start = [n](long long r) { return (uint64_t)r * (2ull * n - (uint64_t)r - 1ull) / 2ull; };
    while (ii > 0 && start(ii) > t) ii--;
    while (ii < (long long)n - 2 && start(ii + 1) <= t) ii++;
    i = (uint32_t)ii;
    j = (uint32_t)(t - start(ii)) + i + 1;
}

}  // namespace gkd
