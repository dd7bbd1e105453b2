// synth.cu -- counter-based synthetic genomes/proteomes (SURVEY section 8d generator).  Bench and
// test utility: position p of descendant `member` of ancestor `family` depends only on
// (seed, family, member, p), so the device kernel and the host loop produce identical bytes.
#include "gkd_internal.cuh"

namespace gkd {

__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__host__ __device__ inline char synth_residue(uint64_t p, uint64_t seed, uint32_t family, uint32_t member,
                                              uint64_t rate_u64, int protein) {
    const uint64_t fam_key = mix64(seed ^ (0xA5A5A5A5ull + family) * 0xD6E8FEB86659FD93ull);
    const uint64_t mem_key = mix64(fam_key ^ (0x5EED0000ull + member) * 0xCA5A826395121157ull);
    const uint32_t radix = protein ? 20u : 4u;
    uint32_t code = (uint32_t)(mix64(fam_key + p) % radix);
    if (member != 0) {
        uint64_t u = mix64(mem_key + p);
        if (u < rate_u64) {  // substitute with one of the other residues
            uint32_t delta = 1u + (uint32_t)(mix64(u ^ mem_key) % (radix - 1u));
            code = (code + delta) % radix;
        }
    }
    return protein ? "ACDEFGHIKLMNPQRSTVWY"[code] : "acgt"[code];
}

inline uint64_t rate_to_u64(double rate) {
    if (rate <= 0) return 0;
    if (rate >= 1) return ~0ull;
    return (uint64_t)(rate * 18446744073709551616.0);
}

__global__ void k_synth(char *dst, uint64_t len, uint64_t seed, uint32_t family, uint32_t member, uint64_t rate_u64,
                        int protein) {
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < len; p += (uint64_t)gridDim.x * blockDim.x)
        dst[p] = synth_residue(p, seed, family, member, rate_u64, protein);
}

cudaError_t launch_synth(char *dst, uint64_t len, uint64_t seed, uint32_t family, uint32_t member, double rate,
                         int protein, cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    uint64_t blocks = (len + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_synth<<<(unsigned)blocks, 256, 0, s>>>(dst, len, seed, family, member, rate_to_u64(rate), protein);
    return cudaGetLastError();
}

void synth_host(char *dst, uint64_t len, uint64_t seed, uint32_t family, uint32_t member, double rate, int protein) {
    uint64_t r = rate_to_u64(rate);
    for (uint64_t p = 0; p < len; p++) dst[p] = synth_residue(p, seed, family, member, r, protein);
}

}  // namespace gkd
