// pack_encode.cu -- kernel 1 (residue text -> packed codes + invalid mask) and kernel 2 (rolling
// canonical k-mer encoding) of libgkd.so.  sm_100a only.
//
// Reference semantics restated:
//   DnaKmers / GenomeKmers: lower-case the sequence, take every K-substring of it and of its reverse
//   complement (KmerType.DNA.createKmers, FastaDistanceProcessor.java:153,184; GenomeProcessor.java:109).
//   Here both strands are represented by ONE canonical key min(fwd, revcomp) per position; the
//   both-strand set sizes are recovered exactly as 2|C| - P (palindromes P counted in kernel 3).
//   ProteinKmers: every K-substring, single strand (ProteinKmerReader.java:100-101); key = the K raw
//   bytes, first character most significant.
//   Kernel 2 emits h = mix(key), the bijectively mixed key the sets are ordered by (gkd_internal.cuh).
// Bound: HBM.  Algorithmic bytes: kernel 1 = 1 B read + 0.375 B written per residue (2-bit code +
// 1-bit mask); kernel 2 = 0.375 B read + 8 B written per k-mer position.
#include "gkd_internal.cuh"

namespace gkd {

// ------------------------------------------------------------------------------------------------
// kernel 1: pack
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t v) {
    // bit 7 of each byte set iff that byte of v is non-zero
    return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}

// one thread = 32 residues = one uint64 of codes + one uint32 of mask
__global__ void __launch_bounds__(256) k_pack_dna(const char *__restrict__ text, uint64_t n_pos, uint64_t n_words,
                                                   uint64_t *__restrict__ codes, uint32_t *__restrict__ mask, int rna) {
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint64_t p0 = w * PACK_POS_PER_WORD;
    uint32_t x[8];
    const char *src = text + p0;
    if (p0 + 32 <= n_pos && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        uint4 a = __ldg(reinterpret_cast<const uint4 *>(src));
        uint4 b = __ldg(reinterpret_cast<const uint4 *>(src) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
        x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                uint64_t p = p0 + q * 4 + e;
                uint32_t c = (p < n_pos) ? (uint8_t)text[p] : 0u;
                v |= c << (8 * e);
            }
            x[q] = v;
        }
    }
    uint64_t code = 0;
    uint32_t inval = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        uint32_t v = x[q];
        uint32_t lower = v | 0x20202020u;  // String.toLowerCase for letters
        if (rna) {                         // u reads as t
            uint32_t is_u = ~nonzero_bytes(lower ^ 0x75757575u) & 0x80808080u;
            lower &= ~(is_u >> 7);
        }
        // a=0 c=1 g=2 t=3 from bits 1..3 of the character
        uint32_t t = ((v >> 1) ^ (v >> 2)) & 0x03030303u;
        // the character each code stands for; a mismatch marks the position invalid
        uint32_t sel = (t & 0x3u) | ((t >> 4) & 0x30u) | ((t >> 8) & 0x300u) | ((t >> 12) & 0x3000u);
        uint32_t expect = __byte_perm(0x74676361u, 0u, sel);
        uint32_t bad = nonzero_bytes(expect ^ lower);
        uint32_t c8 = ((t * 0x00041041u) >> 18) & 0xFFu;                            // 4 codes -> 8 bits
        uint32_t m4 = ((((bad >> 7) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu;    // 4 flags -> 4 bits
        code |= (uint64_t)c8 << (8 * q);
        inval |= m4 << (4 * q);
    }
    // invalid positions carry code 0 so the stream is deterministic
    codes[w] = code;
    mask[w] = inval;
}

__global__ void __launch_bounds__(256) k_pack_prot(const char *__restrict__ text, uint64_t n_pos, uint64_t n_words,
                                                    uint8_t *__restrict__ codes, uint32_t *__restrict__ mask) {
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint64_t p0 = w * PACK_POS_PER_WORD;
    uint32_t inval = 0;
    uint32_t x[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        uint32_t v = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            uint64_t p = p0 + q * 4 + e;
            uint32_t c = (p < n_pos) ? (uint8_t)text[p] : 0u;
            v |= c << (8 * e);
            if (c == (uint8_t)STREAM_SEPARATOR) inval |= 1u << (q * 4 + e);
        }
        x[q] = v;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(codes + p0);  // code buffer is 32-byte aligned and padded
    dst[0] = make_uint4(x[0], x[1], x[2], x[3]);
    dst[1] = make_uint4(x[4], x[5], x[6], x[7]);
    mask[w] = inval;
}

cudaError_t launch_pack_dna(const char *text, uint64_t n_pos, uint64_t *codes, uint32_t *mask, int rna,
                            cudaStream_t s) {
    // one extra all-invalid word so kernel 2 may read word w+1 unconditionally
    uint64_t n_words = (n_pos + PACK_POS_PER_WORD - 1) / PACK_POS_PER_WORD + 1;
    uint64_t blocks = (n_words + 255) / 256;
    k_pack_dna<<<(unsigned)blocks, 256, 0, s>>>(text, n_pos, n_words, codes, mask, rna);
    return cudaGetLastError();
}

cudaError_t launch_pack_prot(const char *text, uint64_t n_pos, uint8_t *codes, uint32_t *mask, cudaStream_t s) {
    uint64_t n_words = (n_pos + PACK_POS_PER_WORD - 1) / PACK_POS_PER_WORD + 1;
    uint64_t blocks = (n_words + 255) / 256;
    k_pack_prot<<<(unsigned)blocks, 256, 0, s>>>(text, n_pos, n_words, codes, mask);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// kernel 2: rolling canonical encode
// ------------------------------------------------------------------------------------------------
constexpr int ENC_STRIDE = ENC_PER_THREAD + 1;  // smem padding: conflict-free 64-bit stores

// The ENC_PER_THREAD k-mer slots that start at stream position my0: emit(j, h) receives h = mix(key) of slot
// my0 + j, or KEY_SENTINEL when the window holds an invalid position / runs past the genome.  j is a
// compile-time index in every call, so callers may keep the keys in registers.
template <int ALPHA, class Emit>
__device__ __forceinline__ void encode_thread(const BatchGenome &G, int k, const MixParams &mix, uint32_t my0, Emit emit) {
    const uint32_t w0 = my0 / PACK_POS_PER_WORD;
    const uint32_t off = my0 % PACK_POS_PER_WORD;  // 0 or 16
    // invalid bits for relative positions 0..47
    const uint64_t inv = (((uint64_t)__ldg(G.mask + w0 + 1) << 32) | __ldg(G.mask + w0)) >> off;
    const uint64_t kbits = (k >= 64) ? ~0ull : ((1ull << k) - 1);
    if (ALPHA == GKD_PROT) {
        const uint64_t kmask = (k >= 8) ? ~0ull : ((1ull << (8 * k)) - 1);
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(G.codes) + my0;
        // 16 slots need bytes [0, 16 + k - 1) <= 23: three aligned 8-byte loads
        const uint64_t *q = reinterpret_cast<const uint64_t *>(bytes);
        const uint64_t b0 = __ldg(q), b1 = __ldg(q + 1), b2 = __ldg(q + 2);
        auto byte_at = [&](int r) { return (uint64_t)((uint32_t)((r < 8 ? b0 : (r < 16 ? b1 : b2)) >> (8 * (r & 7))) & 0xFFu); };
        uint64_t key = 0;
        for (int r = 0; r < k - 1; r++) key = ((key << 8) | byte_at(r)) & kmask;  // first k-1 bytes of slot 0
#pragma unroll
        for (int j = 0; j < ENC_PER_THREAD; j++) {
            key = ((key << 8) | byte_at(j + k - 1)) & kmask;  // key = bytes [j, j + k)
            const bool ok = ((inv >> j) & kbits) == 0 && (my0 + j) < G.n_slots;
            emit(j, ok ? mix_key(key, mix) : KEY_SENTINEL);
        }
    } else {
        const uint64_t kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1);
        const int top = 2 * (k - 1);
        const uint64_t *cw = G.codes + w0;
        const uint64_t c0 = __ldg(cw), c1 = __ldg(cw + 1);
        uint64_t lo, hi;  // 128-bit little-endian window starting at this thread's first base
        if (off) {
            lo = (c0 >> (2 * off)) | (c1 << (64 - 2 * off));
            hi = c1 >> (2 * off);
        } else {
            lo = c0;
            hi = c1;
        }
        // v = the K bases as stored (first base least significant): as a number that is the
        // REVERSED k-mer, so the reverse-complement key is simply its complement.
        uint64_t v = lo & kmask;
        uint64_t fwd = reverse_pairs(v) >> (64 - 2 * k);
#pragma unroll
        for (int j = 0; j < ENC_PER_THREAD; j++) {
            const uint64_t rc = (~v) & kmask;
            const uint64_t key = fwd < rc ? fwd : rc;
            const bool ok = ((inv >> j) & kbits) == 0 && (my0 + j) < G.n_slots;
            emit(j, ok ? mix_key(key, mix) : KEY_SENTINEL);  // sets are kept in mixed-key order
            // slide one base: the 128-bit window shifts right by one code
            lo = (lo >> 2) | (hi << 62);
            hi >>= 2;
            v = lo & kmask;
            fwd = ((fwd << 2) | ((v >> top) & 3ull)) & kmask;
        }
    }
}

template <int ALPHA>
__global__ void __launch_bounds__(ENC_THREADS)
    k_encode(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, int k, MixParams mix,
             uint64_t *__restrict__ keys_out) {
    __shared__ uint64_t stage[ENC_THREADS * ENC_STRIDE];
    __shared__ uint32_t s_g;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const uint32_t tile = blockIdx.x - G.tile_first;
    const uint32_t slot0 = tile * ENC_TILE;                    // first slot of this tile
    const uint32_t my0 = slot0 + threadIdx.x * ENC_PER_THREAD;  // first slot of this thread
    uint64_t *mine = stage + threadIdx.x * ENC_STRIDE;
    if (my0 < G.n_slots) encode_thread<ALPHA>(G, k, mix, my0, [&](int j, uint64_t h) { mine[j] = h; });
    __syncthreads();
    // coalesced write-out of the tile
    const uint32_t remaining = G.n_slots > slot0 ? G.n_slots - slot0 : 0;
    const uint32_t count = remaining < (uint32_t)ENC_TILE ? remaining : (uint32_t)ENC_TILE;
    uint64_t *dst = keys_out + G.raw_off + slot0;
#pragma unroll 4
    for (uint32_t idx = threadIdx.x; idx < count; idx += ENC_THREADS)
        dst[idx] = stage[(idx / ENC_PER_THREAD) * ENC_STRIDE + (idx % ENC_PER_THREAD)];
}

// One global atomic per (tile, bin) reserves the tile's room in every bin; s_bin[b] turns from the tile's key count
// into its offset in bin b (0xFFFFFFFF: the bin would overflow).  A thread owns up to four bins per round and
// issues their atomics back to back, so their latencies overlap instead of adding up.
__device__ __forceinline__ void reserve_bins(uint32_t *s_bin, uint32_t n_bins, uint32_t *__restrict__ cursor, uint32_t cap,
                                             uint32_t *__restrict__ overflow) {
    for (uint32_t b0 = threadIdx.x; b0 < n_bins; b0 += 4u * ENC_THREADS) {
        uint32_t v[4], base[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t b = b0 + u * ENC_THREADS;
            v[u] = b < n_bins ? s_bin[b] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) base[u] = v[u] ? atomicAdd(&cursor[b0 + u * ENC_THREADS], v[u]) : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (v[u]) {
                if (base[u] + v[u] > cap) {
                    atomicExch(overflow, 1u);
                    base[u] = 0xFFFFFFFFu;
                }
                s_bin[b0 + u * ENC_THREADS] = base[u];
            }
    }
}

// ---- kernel 2 fused with the MSD partition of kernel 3 (sort_msd.cu) --------------------------------------
// The mixed keys are uniform, so a genome's bins (top p bits of h) have predictable sizes and get FIXED
// capacity regions: no histogram pass, no scan.  The keys stay in registers; a tile ranks them by bin with
// shared-memory atomics, reserves room in every bin with one global atomic per (tile, bin) and writes each key
// straight to its bin.  A bin that would overflow (heavily repeated k-mers) raises *overflow and the host
// re-runs the batch on the LSD path.
template <int ALPHA>
__global__ void __launch_bounds__(ENC_THREADS, 4)
    k_encode_scatter(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, const MsdGenome *__restrict__ msd, int k,
                     MixParams mix, uint64_t *__restrict__ bins_out, uint32_t *__restrict__ bin_cursor,
                     uint32_t *__restrict__ overflow) {
    extern __shared__ uint32_t s_bin[];  // [2^p] keys of this tile per bin, then the reserved offset in the bin
    __shared__ uint32_t s_g;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const MsdGenome M = msd[s_g];
    const uint32_t n_bins = 1u << M.p;
    for (uint32_t i = threadIdx.x; i < n_bins; i += ENC_THREADS) s_bin[i] = 0;
    __syncthreads();
    const int shift = mix.bits - (int)M.p;
    const uint32_t my0 = (blockIdx.x - G.tile_first) * ENC_TILE + threadIdx.x * ENC_PER_THREAD;
    uint64_t key[ENC_PER_THREAD];
    uint32_t rank[ENC_PER_THREAD];
#pragma unroll
    for (int j = 0; j < ENC_PER_THREAD; j++) key[j] = KEY_SENTINEL, rank[j] = 0;
    if (my0 < G.n_slots)
        encode_thread<ALPHA>(G, k, mix, my0, [&](int j, uint64_t h) {
            key[j] = h;
            if (h != KEY_SENTINEL) rank[j] = atomicAdd(&s_bin[M.p ? (uint32_t)(h >> shift) : 0u], 1u);
        });
    __syncthreads();
    reserve_bins(s_bin, n_bins, bin_cursor + M.bin_first, M.cap, overflow);
    __syncthreads();
    uint64_t *out = bins_out + M.bins_off;
#pragma unroll
    for (int j = 0; j < ENC_PER_THREAD; j++)
        if (key[j] != KEY_SENTINEL) {
            const uint32_t b = M.p ? (uint32_t)(key[j] >> shift) : 0u;
            const uint32_t base = s_bin[b];
            if (base != 0xFFFFFFFFu) out[(uint64_t)b * M.cap + base + rank[j]] = key[j];
        }
}

// the same partition for keys that already exist (imported key arrays): in = h per slot
__global__ void __launch_bounds__(ENC_THREADS)
    k_keys_scatter(const BatchGenome *__restrict__ genomes, uint32_t n_genomes, const MsdGenome *__restrict__ msd, int key_bits,
                   const uint64_t *__restrict__ in, uint64_t *__restrict__ bins_out, uint32_t *__restrict__ bin_cursor,
                   uint32_t *__restrict__ overflow) {
    extern __shared__ uint32_t s_bin[];
    __shared__ uint32_t s_g;
    if (threadIdx.x == 0) s_g = find_genome(genomes, n_genomes, blockIdx.x);
    __syncthreads();
    const BatchGenome G = genomes[s_g];
    const MsdGenome M = msd[s_g];
    const uint32_t n_bins = 1u << M.p;
    for (uint32_t i = threadIdx.x; i < n_bins; i += ENC_THREADS) s_bin[i] = 0;
    __syncthreads();
    const int shift = key_bits - (int)M.p;
    const uint32_t slot0 = (blockIdx.x - G.tile_first) * ENC_TILE;
    const uint32_t count = min((uint32_t)ENC_TILE, G.n_slots - slot0);
    const uint64_t *src = in + G.raw_off + slot0;
    uint64_t key[ENC_PER_THREAD];
    uint32_t rank[ENC_PER_THREAD];
#pragma unroll
    for (int j = 0; j < ENC_PER_THREAD; j++) {
        const uint32_t idx = j * ENC_THREADS + threadIdx.x;
        key[j] = idx < count ? src[idx] : KEY_SENTINEL;
        rank[j] = 0;
        if (key[j] != KEY_SENTINEL) rank[j] = atomicAdd(&s_bin[M.p ? (uint32_t)(key[j] >> shift) : 0u], 1u);
    }
    __syncthreads();
    reserve_bins(s_bin, n_bins, bin_cursor + M.bin_first, M.cap, overflow);
    __syncthreads();
    uint64_t *out = bins_out + M.bins_off;
#pragma unroll
    for (int j = 0; j < ENC_PER_THREAD; j++)
        if (key[j] != KEY_SENTINEL) {
            const uint32_t b = M.p ? (uint32_t)(key[j] >> shift) : 0u;
            const uint32_t base = s_bin[b];
            if (base != 0xFFFFFFFFu) out[(uint64_t)b * M.cap + base + rank[j]] = key[j];
        }
}

cudaError_t launch_encode(const BatchGenome *genomes, uint32_t n_genomes, uint32_t n_tiles, int alphabet, int k,
                          MixParams mix, uint64_t *keys_out, cudaStream_t s) {
    if (n_tiles == 0) return cudaSuccess;
    if (alphabet == GKD_PROT) k_encode<GKD_PROT><<<n_tiles, ENC_THREADS, 0, s>>>(genomes, n_genomes, k, mix, keys_out);
    else k_encode<GKD_DNA><<<n_tiles, ENC_THREADS, 0, s>>>(genomes, n_genomes, k, mix, keys_out);
    return cudaGetLastError();
}

}  // namespace gkd

namespace gkd {

cudaError_t launch_encode_scatter(const BatchGenome *genomes, uint32_t n_genomes, uint32_t n_tiles, const MsdGenome *msd,
                                  uint32_t max_p, int alphabet, int k, MixParams mix, const uint64_t *keys_in, uint64_t *bins_out,
                                  uint32_t *bin_cursor, uint32_t *overflow, cudaStream_t s) {
    if (n_tiles == 0) return cudaSuccess;
    const uint32_t smem = (1u << max_p) * 4;
    if (keys_in)
        k_keys_scatter<<<n_tiles, ENC_THREADS, smem, s>>>(genomes, n_genomes, msd, mix.bits, keys_in, bins_out, bin_cursor, overflow);
    else if (alphabet == GKD_PROT)
        k_encode_scatter<GKD_PROT><<<n_tiles, ENC_THREADS, smem, s>>>(genomes, n_genomes, msd, k, mix, bins_out, bin_cursor, overflow);
    else
        k_encode_scatter<GKD_DNA><<<n_tiles, ENC_THREADS, smem, s>>>(genomes, n_genomes, msd, k, mix, bins_out, bin_cursor, overflow);
    return cudaGetLastError();
}

}  // namespace gkd
