// sketch.cu -- MinHash sketches from the exact sets.
//
// Reference call sites (the implementing classes are external and the hash is UNPINNED here):
//   SequenceKmers.hashSet(width)           SketchProcessor.java:91, WidthProcessor.java:177, MashProcessor.java:116-150
//   new Sketch(int[] signature, name), Sketch.distance(other)      WidthProcessor.java:183
// Restated contract: hashSet(width) = the `width` smallest DISTINCT hash codes of the set's k-mer strings in
// ascending (signed Java int) order -- fewer when the set has fewer distinct codes ("dwarves",
// WidthProcessor.java:178).  The hash of a string is a switch: java.lang.String.hashCode (exact k-mer sets
// are HashSet<String>, so that code is at hand upstream) or murmur3_x86_32 with seed 0 (build.xml:30 ships
// com.github.eprst:murmur3 next to the sequence module).  For nucleotide sets the members are the
// lower-case strings of BOTH strands, so each canonical key contributes the codes of the k-mer and of its
// reverse complement.  Sketch.distance is the bottom-w estimator: walk the union of the two signatures in
// ascending order for w = min(|A|, |B|) steps, count the codes present in both, distance = 1 - m / w
// (1.0 when either signature is empty).
//
// Device plan: one streaming pass over the set (un-mix, rebuild the characters, hash) keeps the codes below a
// threshold chosen from the set size; one CTA sorts those few thousand candidates in shared memory (bitonic),
// drops duplicates and writes the first `width`.  Pair distances: one thread per pair merges two signatures.
#include "gkd_internal.cuh"

namespace gkd {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// murmur3_x86_32 over k bytes produced by byte_at(i), seed 0
template <class F>
__device__ __forceinline__ uint32_t murmur3_32(F byte_at, int k) {
    uint32_t h = 0;
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    int i = 0;
    for (; i + 4 <= k; i += 4) {
        uint32_t kk = byte_at(i) | (byte_at(i + 1) << 8) | (byte_at(i + 2) << 16) | (byte_at(i + 3) << 24);
        kk *= c1;
        kk = rotl32(kk, 15);
        kk *= c2;
        h ^= kk;
        h = rotl32(h, 13);
        h = h * 5u + 0xe6546b64u;
    }
    uint32_t kk = 0;
    const int rem = k - i;
    if (rem >= 3) kk ^= byte_at(i + 2) << 16;
    if (rem >= 2) kk ^= byte_at(i + 1) << 8;
    if (rem >= 1) {
        kk ^= byte_at(i);
        kk *= c1;
        kk = rotl32(kk, 15);
        kk *= c2;
        h ^= kk;
    }
    h ^= (uint32_t)k;
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

template <class F>
__device__ __forceinline__ uint32_t string_hash(F byte_at, int k, int kind) {
    if (kind == GKD_HASH_MURMUR3) return murmur3_32(byte_at, k);
    uint32_t h = 0;  // java.lang.String.hashCode
    for (int i = 0; i < k; i++) h = 31u * h + byte_at(i);
    return h;
}

// One thread per bucket: rebuild every key of the bucket, hash its string(s), keep the codes whose biased
// value (code ^ 0x80000000: unsigned order == signed order) is <= thresh.
template <typename LowT>
__global__ void __launch_bounds__(256)
    k_sketch_filter(SubSet S, MixParams mix, int alphabet, int k, int both, int kind, uint32_t thresh,
                    uint32_t *__restrict__ cand, uint32_t cap, uint32_t *__restrict__ n_cand) {
    const LowT *lows = (const LowT *)S.lows;
    const uint32_t n_buckets = 1u << S.level;
    const int rest = mix.bits - (int)S.level;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_buckets; b += gridDim.x * blockDim.x) {
        const uint32_t lo = S.offs[b], hi = S.offs[b + 1];
        for (uint32_t i = lo; i < hi; i++) {
            uint64_t h = (uint64_t)lows[i];
            if (sizeof(LowT) == 4 && S.level > 0) h = ((uint64_t)b << rest) | (h & ((1ull << rest) - 1ull));
            const uint64_t key = unmix_key(h, mix);
            uint32_t codes[2];
            int n_codes = 1;
            if (alphabet == GKD_PROT) {
                codes[0] = string_hash([&](int p) { return (uint32_t)(key >> (8 * (k - 1 - p))) & 0xFFu; }, k, kind);
            } else {
                const char letters[4] = {'a', 'c', 'g', 't'};
                codes[0] = string_hash([&](int p) { return (uint32_t)letters[(key >> (2 * (k - 1 - p))) & 3u]; }, k, kind);
                if (both) {  // reverse complement strand: base p is the complement of base k-1-p
                    codes[1] = string_hash([&](int p) { return (uint32_t)letters[3u - ((key >> (2 * p)) & 3u)]; }, k, kind);
                    n_codes = 2;
                }
            }
            for (int c = 0; c < n_codes; c++) {
                const uint32_t u = codes[c] ^ 0x80000000u;
                if (u <= thresh) {
                    const uint32_t at = atomicAdd(n_cand, 1u);
                    if (at < cap) cand[at] = u;
                }
            }
        }
    }
}

// One CTA: bitonic sort of n <= P candidates in shared memory, drop duplicates, write the first `width`
// as signed Java ints; *n_out = number written, *n_distinct = distinct candidates seen.
__global__ void __launch_bounds__(1024)
    k_sketch_finish(const uint32_t *__restrict__ cand, uint32_t n, uint32_t P, uint32_t width, int32_t *__restrict__ out,
                    uint32_t *__restrict__ n_out, uint32_t *__restrict__ n_distinct) {
    extern __shared__ uint32_t s_v[];
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) s_v[i] = i < n ? cand[i] : 0xFFFFFFFFu;
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < P / 2; i += blockDim.x) {
                const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint32_t a = s_v[lo], b = s_v[hi];
                if ((a > b) == up) {
                    s_v[lo] = b;
                    s_v[hi] = a;
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {  // a few thousand elements: a serial compaction is cheaper than a scan here
        uint32_t w = 0, d = 0;
        for (uint32_t i = 0; i < n; i++) {
            if (i > 0 && s_v[i] == s_v[i - 1]) continue;
            if (w < width) out[w++] = (int32_t)(s_v[i] ^ 0x80000000u);
            d++;
        }
        *n_out = w;
        *n_distinct = d;
    }
}

cudaError_t launch_sketch_filter(SubSet set, MixParams mix, int low_bits, int alphabet, int k, int both, int kind,
                                 uint32_t thresh, uint32_t *cand, uint32_t cap, uint32_t *n_cand, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(n_cand, 0, 4, s);
    if (e != cudaSuccess) return e;
    if (set.n == 0) return cudaSuccess;
    uint32_t blocks = ((1u << set.level) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (low_bits == 32)
        k_sketch_filter<uint32_t><<<blocks, 256, 0, s>>>(set, mix, alphabet, k, both, kind, thresh, cand, cap, n_cand);
    else
        k_sketch_filter<uint64_t><<<blocks, 256, 0, s>>>(set, mix, alphabet, k, both, kind, thresh, cand, cap, n_cand);
    return cudaGetLastError();
}

cudaError_t sketch_configure() {
    return cudaFuncSetAttribute(k_sketch_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SKETCH_CAP * 4));
}

cudaError_t launch_sketch_finish(const uint32_t *cand, uint32_t n, uint32_t width, int32_t *out, uint32_t *n_out,
                                 uint32_t *n_distinct, cudaStream_t s) {
    uint32_t P = 2;
    while (P < n) P <<= 1;
    if (P > SKETCH_CAP) return cudaErrorInvalidValue;
    k_sketch_finish<<<1, 1024, P * 4, s>>>(cand, n, P, width, out, n_out, n_distinct);
    return cudaGetLastError();
}

// Sketch.distance for a list of pairs: signatures are rows of `sig` (stride `width`, lengths in `len`)
__global__ void __launch_bounds__(128)
    k_sketch_distance(const int32_t *__restrict__ sig, const uint32_t *__restrict__ len, uint32_t width,
                      const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint64_t n_pairs,
                      double *__restrict__ dist) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const int32_t *sa = sig + (size_t)a[t] * width, *sb = sig + (size_t)b[t] * width;
    const uint32_t na = len[a[t]], nb = len[b[t]];
    const uint32_t w = na < nb ? na : nb;
    uint32_t i = 0, j = 0, m = 0;
    for (uint32_t step = 0; step < w; step++) {  // i, j < w <= na, nb throughout
        const int32_t x = sa[i], y = sb[j];
        m += (x == y) ? 1u : 0u;
        i += (x <= y) ? 1u : 0u;
        j += (y <= x) ? 1u : 0u;
        if (i >= na || j >= nb) break;
    }
    dist[t] = w == 0 ? 1.0 : 1.0 - (double)m / (double)w;
}

cudaError_t launch_sketch_distance(const int32_t *sig, const uint32_t *len, uint32_t width, const uint32_t *a,
                                   const uint32_t *b, uint64_t n_pairs, double *dist, cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    k_sketch_distance<<<(unsigned)((n_pairs + 127) / 128), 128, 0, s>>>(sig, len, width, a, b, n_pairs, dist);
    return cudaGetLastError();
}

}  // namespace gkd
