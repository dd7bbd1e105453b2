// gkd_api.cu -- context, memory plan and the extern "C" ABI of libgkd.so (see include/gkd.h).
// One context = one CUDA device, one stream.  No CPU fallback: without a usable device every
// computing entry point fails with GKD_ECUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gkd_internal.cuh"

using namespace gkd;

// host FASTA scanner (fasta.cpp)
struct FastaPiece {
    const char *ptr;
    uint64_t len;
};
struct FastaRecord {
    std::string label, comment;
    std::vector<FastaPiece> lines;  // sequence lines, to be concatenated
};
int gkd_parse_fasta_file(const char *path, std::vector<char> &storage, std::vector<FastaRecord> &records,
                         std::string &err);

namespace {

thread_local std::string g_create_error;

constexpr uint64_t PACK_SLAB_BYTES = 256ull << 20;
constexpr uint64_t STAGE_BYTES = 32ull << 20;  // text staged per piece (multiple of 32 positions)
constexpr int N_STAGE = 3;
constexpr uint64_t DEFAULT_WORKSPACE = 8ull << 30;
constexpr uint64_t PAIR_CHUNK = 64ull << 20;  // pairs per distance launch

struct Part {  // a run of residue text; new_contig starts a new contig (k-mers do not span contigs)
    const char *ptr;
    uint64_t len;
    bool new_contig;
};

struct GenomeRec {
    std::string label, comment;
    uint64_t n_pos = 0;         // stream positions (residues + separators)
    void *d_codes = nullptr;    // packed stream (slab memory)
    uint32_t *d_mask = nullptr;
    bool built = false;
    SetDesc desc{nullptr, 0, 0, nullptr};
};

struct Slab {
    char *base;
    uint64_t size, used;
};

struct DevBuf {  // grow-only scratch buffer
    void *p = nullptr;
    uint64_t cap = 0;
};

}  // namespace

struct gkd_ctx {
    gkd_config cfg{};
    int k = 0;
    int n_sms = 148;
    cudaStream_t stream = nullptr;
    bool poisoned = false;
    std::string err;

    std::vector<GenomeRec> genomes;
    uint32_t built_upto = 0;

    std::vector<Slab> slabs;
    std::vector<std::pair<void *, uint64_t>> set_arenas;   // live set arenas (ptr, bytes)
    std::vector<uint32_t> arena_first_id;                   // first set id stored in each live arena
    std::vector<std::pair<void *, uint64_t>> free_arenas;  // arenas released by gkd_reset, reused best-fit

    char *bounce[N_STAGE] = {nullptr, nullptr, nullptr};
    cudaEvent_t bounce_ev[N_STAGE] = {nullptr, nullptr, nullptr};
    char *stage_dev[N_STAGE] = {nullptr, nullptr, nullptr};
    int stage_next = 0;

    DevBuf keys_a, keys_b, tile_hist, tile_uniq, genome_counts, batch_genomes, uniq_dst;
    DevBuf d_sets, counts, pal_counts, d_inter, d_dist, ids_a, ids_b, work_counter;
    bool sets_dirty = true;

    cudaEvent_t ev[8] = {};
    gkd_metrics m{};
};

namespace {

int fail(gkd_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            c->poisoned = true;                                                                          \
            return fail(c, e__ == cudaErrorMemoryAllocation ? GKD_ENOMEM : GKD_ECUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                    \
        }                                                                                                \
    } while (0)

#define CHECK_CTX(c)                                                                   \
    do {                                                                               \
        if (!(c)) return GKD_EINVAL;                                                   \
        if ((c)->poisoned) return fail((c), GKD_ECUDA, "context poisoned by an earlier CUDA error: %s", (c)->err.c_str()); \
    } while (0)

int ensure(gkd_ctx *c, DevBuf &b, uint64_t bytes) {
    if (bytes <= b.cap) return GKD_OK;
    if (b.p) CK(cudaFreeAsync(b.p, c->stream));
    b.p = nullptr;
    b.cap = 0;
    uint64_t want = bytes + bytes / 8 + 256;
    CK(cudaMallocAsync(&b.p, want, c->stream));
    b.cap = want;
    return GKD_OK;
}

int slab_alloc(gkd_ctx *c, uint64_t bytes, void **out) {
    bytes = (bytes + 255) & ~255ull;
    for (auto &s : c->slabs) {
        if (s.size - s.used >= bytes) {
            *out = s.base + s.used;
            s.used += bytes;
            return GKD_OK;
        }
    }
    Slab s;
    s.size = std::max<uint64_t>(PACK_SLAB_BYTES, bytes);
    s.used = bytes;
    void *p = nullptr;
    CK(cudaMallocAsync(&p, s.size, c->stream));
    s.base = (char *)p;
    c->slabs.push_back(s);
    *out = s.base;
    return GKD_OK;
}

enum MemKind { MEM_PAGEABLE = 0, MEM_PINNED = 1, MEM_DEVICE = 2 };

MemKind classify(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return MEM_PAGEABLE;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return MEM_DEVICE;
    if (a.type == cudaMemoryTypeHost) return MEM_PINNED;
    return MEM_PAGEABLE;
}

double elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return ms;
}

// ---- ingest -------------------------------------------------------------------------------------------
int add_genome(gkd_ctx *c, const std::vector<Part> &parts, const std::string &label, const std::string &comment,
               uint32_t *out_id) {
    const bool prot = c->cfg.alphabet == GKD_PROT;
    uint64_t n_pos = 0;
    bool first = true;
    for (auto &p : parts) {
        if (p.new_contig && !first) n_pos += 1;  // separator position between contigs
        n_pos += p.len;
        first = false;
    }
    if (n_pos >= 0xFFFF0000ull) return fail(c, GKD_EINVAL, "sequence of %llu residues exceeds the 2^32 position limit", (unsigned long long)n_pos);
    GenomeRec g;
    g.label = label;
    g.comment = comment;
    g.n_pos = n_pos;
    const uint64_t n_words = (n_pos + PACK_POS_PER_WORD - 1) / PACK_POS_PER_WORD + 2;
    void *codes = nullptr, *mask = nullptr;
    int rc = slab_alloc(c, n_words * (prot ? 32 : 8), &codes);
    if (rc) return rc;
    rc = slab_alloc(c, n_words * 4, &mask);
    if (rc) return rc;
    g.d_codes = codes;
    g.d_mask = (uint32_t *)mask;

    // memory kind of the inputs (all parts of one call are assumed to live in the same kind)
    MemKind kind = MEM_PAGEABLE;
    if (!parts.empty()) {
        kind = classify(parts.front().ptr);
        if (parts.size() > 1 && classify(parts.back().ptr) != kind)
            return fail(c, GKD_EINVAL, "all sequence pieces of one call must be in the same kind of memory");
    }
    // walk the stream in pieces of at most STAGE_BYTES positions
    size_t pi = 0;          // current part
    uint64_t pofs = 0;      // offset inside current part
    bool pending_sep = false;
    uint64_t q0 = 0;
    do {
        const uint64_t q1 = std::min<uint64_t>(n_pos, q0 + STAGE_BYTES);
        const uint64_t plen = q1 - q0;
        const int s = c->stage_next;
        c->stage_next = (c->stage_next + 1) % N_STAGE;
        char *dev = c->stage_dev[s];
        if (kind == MEM_PAGEABLE) {
            CK(cudaEventSynchronize(c->bounce_ev[s]));
            char *dst = c->bounce[s];
            uint64_t w = 0;
            while (w < plen) {
                if (pending_sep) {
                    dst[w++] = STREAM_SEPARATOR;
                    pending_sep = false;
                    continue;
                }
                const Part &p = parts[pi];
                uint64_t take = std::min<uint64_t>(p.len - pofs, plen - w);
                memcpy(dst + w, p.ptr + pofs, take);
                w += take;
                pofs += take;
                if (pofs == p.len) {
                    pi++;
                    pofs = 0;
                    if (pi < parts.size() && parts[pi].new_contig) pending_sep = true;
                }
            }
            if (plen) CK(cudaMemcpyAsync(dev, dst, plen, cudaMemcpyHostToDevice, c->stream));
            CK(cudaEventRecord(c->bounce_ev[s], c->stream));
            c->m.h2d_bytes += plen;
        } else {
            // pinned host or device memory: copy each run straight into the staged text; the
            // separators are the zero fill
            if (plen) CK(cudaMemsetAsync(dev, STREAM_SEPARATOR, plen, c->stream));
            uint64_t w = 0;
            while (w < plen) {
                if (pending_sep) {
                    w++;
                    pending_sep = false;
                    continue;
                }
                const Part &p = parts[pi];
                uint64_t take = std::min<uint64_t>(p.len - pofs, plen - w);
                if (take) CK(cudaMemcpyAsync(dev + w, p.ptr + pofs, take, cudaMemcpyDefault, c->stream));
                w += take;
                pofs += take;
                if (pofs == p.len) {
                    pi++;
                    pofs = 0;
                    if (pi < parts.size() && parts[pi].new_contig) pending_sep = true;
                }
            }
            if (kind == MEM_PINNED) c->m.h2d_bytes += plen;
        }
        // kernel 1 on this piece (q0 is a multiple of 32 positions)
        if (prot)
            CK(launch_pack_prot(dev, plen, (uint8_t *)codes + q0, (uint32_t *)mask + q0 / 32, c->stream));
        else
            CK(launch_pack_dna(dev, plen, (uint64_t *)codes + q0 / 32, (uint32_t *)mask + q0 / 32,
                               c->cfg.alphabet == GKD_RNA, c->stream));
        c->m.launches++;
        q0 = q1;
    } while (q0 < n_pos);
    // contract (gkd.h): inputs are consumed before the call returns.  Pageable text was copied into the
    // bounce buffers synchronously; pinned/device text is read by stream-ordered copies, so wait for them.
    if (kind != MEM_PAGEABLE) CK(cudaStreamSynchronize(c->stream));
    c->m.residues_packed += n_pos;
    if (out_id) *out_id = (uint32_t)c->genomes.size();
    c->genomes.push_back(std::move(g));
    return GKD_OK;
}

// ---- set construction ------------------------------------------------------------------------------------
struct BatchItem {
    uint32_t id;        // genome index, or UINT32_MAX for an imported key array
    uint32_t n_slots;
};

// unique/compact the sorted slots of a batch into a fresh set arena and record the descriptors
int finish_batch(gkd_ctx *c, const std::vector<BatchGenome> &bg, const std::vector<uint32_t> &ids, const SortPlan &plan,
                 const uint64_t *sorted) {
    const uint32_t n = (uint32_t)bg.size();
    CK(launch_unique_count((const BatchGenome *)c->batch_genomes.p, plan, sorted, c->cfg.alphabet, c->k, c->stream));
    c->m.launches += 2;
    std::vector<uint64_t> counts(n);
    CK(cudaMemcpyAsync(counts.data(), plan.genome_counts, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    uint64_t arena_keys = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t nu = (uint32_t)counts[i], np = (uint32_t)(counts[i] >> 32);
        arena_keys += set_padded(nu);
        if (np) arena_keys += set_padded(np);
    }
    void *arena = nullptr;
    {
        // reuse an arena released by gkd_reset (best fit) so repeated runs do not make the pool remap
        const uint64_t need = arena_keys * 8 + 256;
        int best = -1;
        for (size_t i = 0; i < c->free_arenas.size(); i++)
            if (c->free_arenas[i].second >= need && (best < 0 || c->free_arenas[i].second < c->free_arenas[best].second))
                best = (int)i;
        if (best >= 0) {
            arena = c->free_arenas[best].first;
            c->set_arenas.push_back(c->free_arenas[best]);
            c->free_arenas.erase(c->free_arenas.begin() + best);
        } else {
            CK(cudaMallocAsync(&arena, need, c->stream));
            c->set_arenas.push_back({arena, need});
        }
        c->arena_first_id.push_back(ids.empty() ? 0u : ids.front());
    }
    std::vector<UniqueDst> dst(n);
    uint64_t *cur = (uint64_t *)arena;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t nu = (uint32_t)counts[i], np = (uint32_t)(counts[i] >> 32);
        dst[i].keys = cur;
        cur += set_padded(nu);
        dst[i].pal_keys = nullptr;
        if (np) {
            dst[i].pal_keys = cur;
            cur += set_padded(np);
        }
        GenomeRec &g = c->genomes[ids[i]];
        g.desc.keys = dst[i].keys;
        g.desc.n = nu;
        g.desc.n_pal = np;
        g.desc.pal_keys = dst[i].pal_keys;
        g.built = true;
        c->m.keys_unique += nu;
    }
    int rc = ensure(c, c->uniq_dst, n * sizeof(UniqueDst));
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->uniq_dst.p, dst.data(), n * sizeof(UniqueDst), cudaMemcpyHostToDevice, c->stream));
    CK(launch_unique_write((const BatchGenome *)c->batch_genomes.p, plan, sorted, (const UniqueDst *)c->uniq_dst.p,
                           c->cfg.alphabet, c->k, c->stream));
    c->m.launches += 2;
    // the host vectors above are pageable: make sure the copies were consumed before they die
    CK(cudaStreamSynchronize(c->stream));
    c->sets_dirty = true;
    return GKD_OK;
}

int plan_batch(gkd_ctx *c, std::vector<BatchGenome> &bg, SortPlan &plan, uint64_t raw_keys, uint32_t n_tiles) {
    int rc;
    if ((rc = ensure(c, c->keys_a, raw_keys * 8))) return rc;
    if ((rc = ensure(c, c->keys_b, raw_keys * 8))) return rc;
    if ((rc = ensure(c, c->tile_hist, (uint64_t)std::max(n_tiles, 1u) * RADIX_BINS * 4))) return rc;
    if ((rc = ensure(c, c->tile_uniq, (uint64_t)std::max(n_tiles, 1u) * 8))) return rc;
    if ((rc = ensure(c, c->genome_counts, bg.size() * 8))) return rc;
    if ((rc = ensure(c, c->batch_genomes, bg.size() * sizeof(BatchGenome)))) return rc;
    CK(cudaMemcpyAsync(c->batch_genomes.p, bg.data(), bg.size() * sizeof(BatchGenome), cudaMemcpyHostToDevice, c->stream));
    plan.n_genomes = (uint32_t)bg.size();
    plan.n_tiles = n_tiles;
    plan.key_bits = c->cfg.alphabet == GKD_PROT ? 8 * c->k : 2 * c->k;
    plan.keys_a = (uint64_t *)c->keys_a.p;
    plan.keys_b = (uint64_t *)c->keys_b.p;
    plan.tile_hist = (uint32_t *)c->tile_hist.p;
    plan.tile_uniq = (uint64_t *)c->tile_uniq.p;
    plan.genome_counts = (uint64_t *)c->genome_counts.p;
    return GKD_OK;
}

int build_batch(gkd_ctx *c, uint32_t first, uint32_t last) {
    std::vector<BatchGenome> bg;
    std::vector<uint32_t> ids;
    uint64_t raw = 0;
    uint32_t tiles = 0;
    for (uint32_t id = first; id < last; id++) {
        GenomeRec &g = c->genomes[id];
        BatchGenome b;
        b.codes = (const uint64_t *)g.d_codes;
        b.mask = g.d_mask;
        b.n_pos = (uint32_t)g.n_pos;
        b.n_slots = g.n_pos >= (uint64_t)c->k ? (uint32_t)(g.n_pos - c->k + 1) : 0;
        b.raw_off = raw;
        b.tile_first = tiles;
        b.n_tiles = (b.n_slots + SORT_TILE - 1) / SORT_TILE;
        raw += ((uint64_t)b.n_slots + 15) & ~15ull;
        tiles += b.n_tiles;
        bg.push_back(b);
        ids.push_back(id);
        c->m.kmer_positions += b.n_slots;
    }
    SortPlan plan{};
    int rc = plan_batch(c, bg, plan, std::max<uint64_t>(raw, 16), tiles);
    if (rc) return rc;
    CK(cudaEventRecord(c->ev[0], c->stream));
    CK(launch_encode((const BatchGenome *)c->batch_genomes.p, plan.n_genomes, tiles, c->cfg.alphabet, c->k, plan.keys_a,
                     c->stream));
    if (tiles) c->m.launches++;
    CK(cudaEventRecord(c->ev[1], c->stream));
    uint64_t *sorted = nullptr;
    uint32_t passes = 0;
    CK(launch_sort((const BatchGenome *)c->batch_genomes.p, plan, &sorted, &passes, c->stream));
    c->m.launches += 3ull * passes;
    c->m.sort_passes = passes;
    c->m.keys_sorted += raw;
    CK(cudaEventRecord(c->ev[2], c->stream));
    rc = finish_batch(c, bg, ids, plan, sorted);
    if (rc) return rc;
    CK(cudaEventRecord(c->ev[3], c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->m.encode_ms += elapsed(c->ev[0], c->ev[1]);
    c->m.sort_ms += elapsed(c->ev[1], c->ev[2]);
    c->m.unique_ms += elapsed(c->ev[2], c->ev[3]);
    return GKD_OK;
}

int upload_sets(gkd_ctx *c) {
    if (!c->sets_dirty) return GKD_OK;
    const size_t n = c->genomes.size();
    std::vector<SetDesc> h(n);
    for (size_t i = 0; i < n; i++) h[i] = c->genomes[i].desc;
    int rc = ensure(c, c->d_sets, std::max<size_t>(n, 1) * sizeof(SetDesc));
    if (rc) return rc;
    if (n) CK(cudaMemcpyAsync(c->d_sets.p, h.data(), n * sizeof(SetDesc), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->sets_dirty = false;
    return GKD_OK;
}

// ---- distances ------------------------------------------------------------------------------------------
// host copy of the pair source (ids live on the host here)
struct HostPairs {
    int mode;
    uint32_t n;
    uint64_t first, count;
    const uint32_t *a, *b;
    uint32_t na, nb;  // lengths of a / b arrays (LIST: count each; RECT: nq, nr)
};

int check_built(gkd_ctx *c, uint32_t id) {
    if (id >= c->genomes.size()) return fail(c, GKD_EINVAL, "set id %u out of range (have %zu)", id, c->genomes.size());
    if (!c->genomes[id].built) return fail(c, GKD_ESTATE, "set %u has not been built; call gkd_build_sets first", id);
    return GKD_OK;
}

int run_pairs(gkd_ctx *c, const HostPairs &hp, uint64_t *inter, double *dist) {
    int rc = upload_sets(c);
    if (rc) return rc;
    const bool nuc = c->cfg.alphabet != GKD_PROT;
    const bool both = nuc && c->cfg.strand_mode == GKD_STRAND_BOTH;
    const bool need_pal = both && (c->k % 2 == 0);
    c->m.pairs = hp.count;
    c->m.intersect_bytes = 0;
    c->m.intersect_ms = c->m.epilogue_ms = 0;

    // validate ids, gather sizes for the segmenting decision and the byte accounting
    uint64_t max_n = 0, min_n = UINT64_MAX;
    long double sum_bytes = 0;
    auto size_of = [&](uint32_t id) -> uint64_t { return c->genomes[id].desc.n; };
    if (hp.mode == PAIRS_UPPER) {
        for (uint32_t i = 0; i < hp.n; i++) {
            if ((rc = check_built(c, i))) return rc;
            max_n = std::max<uint64_t>(max_n, size_of(i));
            if (size_of(i)) min_n = std::min<uint64_t>(min_n, size_of(i));
        }
        // bytes of the requested range, row by row
        uint32_t i0 = 0, j0 = 1;
        if (hp.count) upper_pair(hp.first, hp.n, i0, j0);
        uint64_t left = hp.count;
        std::vector<uint64_t> prefix(hp.n + 1, 0);
        for (uint32_t i = 0; i < hp.n; i++) prefix[i + 1] = prefix[i] + size_of(i);
        for (uint32_t i = i0, j = j0; left > 0 && i + 1 < hp.n; i++, j = i + 1) {
            uint64_t in_row = std::min<uint64_t>(left, hp.n - j);
            sum_bytes += 8.0L * ((long double)in_row * size_of(i) + (long double)(prefix[j + in_row] - prefix[j]));
            left -= in_row;
        }
    } else if (hp.mode == PAIRS_RECT) {
        uint64_t sq = 0, sr = 0;
        for (uint32_t i = 0; i < hp.na; i++) {
            if ((rc = check_built(c, hp.a[i]))) return rc;
            sq += size_of(hp.a[i]);
            max_n = std::max<uint64_t>(max_n, size_of(hp.a[i]));
            if (size_of(hp.a[i])) min_n = std::min<uint64_t>(min_n, size_of(hp.a[i]));
        }
        for (uint32_t i = 0; i < hp.nb; i++) {
            if ((rc = check_built(c, hp.b[i]))) return rc;
            sr += size_of(hp.b[i]);
            max_n = std::max<uint64_t>(max_n, size_of(hp.b[i]));
            if (size_of(hp.b[i])) min_n = std::min<uint64_t>(min_n, size_of(hp.b[i]));
        }
        sum_bytes = 8.0L * ((long double)sq * hp.nb + (long double)sr * hp.na);
    } else {
        for (uint64_t t = 0; t < hp.count; t++) {
            if ((rc = check_built(c, hp.a[t]))) return rc;
            if ((rc = check_built(c, hp.b[t]))) return rc;
            uint64_t s = size_of(hp.a[t]) + size_of(hp.b[t]);
            sum_bytes += 8.0L * s;
            max_n = std::max<uint64_t>(max_n, std::max(size_of(hp.a[t]), size_of(hp.b[t])));
            if (size_of(hp.a[t])) min_n = std::min<uint64_t>(min_n, size_of(hp.a[t]));
            if (size_of(hp.b[t])) min_n = std::min<uint64_t>(min_n, size_of(hp.b[t]));
        }
    }
    c->m.intersect_bytes = (uint64_t)sum_bytes;
    if (hp.count == 0) return GKD_OK;

    // palindrome side lists are tiny; the largest one decides which kernel intersects them
    uint64_t max_pal = 0;
    if (need_pal)
        for (auto &g : c->genomes) max_pal = std::max<uint64_t>(max_pal, g.desc.n_pal);
    const int algo = intersect_select(min_n == UINT64_MAX ? 0 : min_n, max_n, nuc ? 2 * c->k : 8 * c->k);
    const bool small_main = c->cfg.segment_keys == 0 && max_n <= intersect_small_max_keys();
    const bool small_pal = max_pal <= intersect_small_max_keys();
    c->m.intersect_kernel = small_main ? 2u : (uint32_t)algo;

    // merge-path segmenting: whole pairs when there are enough of them to fill the machine
    const uint64_t max_l = std::max<uint64_t>(2 * max_n, 1);
    const uint64_t target_items = (uint64_t)c->n_sms * intersect_items_per_sm(algo);
    uint64_t seg = c->cfg.segment_keys;
    if (seg == 0) {
        if (hp.count >= target_items) seg = max_l;
        else {
            uint64_t total_keys = (uint64_t)(sum_bytes / 8.0L);
            // a segment must amortise its two global diagonal searches: ~5 CTA rounds, or ~70 warp steps
            seg = std::max<uint64_t>(total_keys / target_items, algo ? 4096 : 16384);
            seg = std::min<uint64_t>(seg, max_l);
        }
    }
    seg = std::max<uint64_t>(seg, (uint64_t)intersect_min_segment(algo));
    seg = std::min<uint64_t>(seg, 0xFFFF0000ull);
    const uint32_t max_segs = (uint32_t)((max_l + seg - 1) / seg);

    // device id arrays
    PairSource src{};
    src.mode = hp.mode;
    src.n = hp.n;
    if (hp.mode != PAIRS_UPPER) {
        if ((rc = ensure(c, c->ids_a, (uint64_t)hp.na * 4))) return rc;
        if ((rc = ensure(c, c->ids_b, (uint64_t)hp.nb * 4))) return rc;
        CK(cudaMemcpyAsync(c->ids_a.p, hp.a, (uint64_t)hp.na * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->ids_b.p, hp.b, (uint64_t)hp.nb * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    if ((rc = ensure(c, c->work_counter, 8))) return rc;

    for (uint64_t done = 0; done < hp.count; done += PAIR_CHUNK) {
        const uint64_t cnt = std::min<uint64_t>(PAIR_CHUNK, hp.count - done);
        src.count = cnt;
        if (hp.mode == PAIRS_UPPER) {
            src.first = hp.first + done;
            src.a = src.b = nullptr;
        } else if (hp.mode == PAIRS_RECT) {
            // gkd_query_vs_ref splits by query rows, so a rectangle always fits one launch
            src.first = 0;
            src.a = (const uint32_t *)c->ids_a.p;
            src.b = (const uint32_t *)c->ids_b.p;
            if (hp.count > PAIR_CHUNK) return fail(c, GKD_EINVAL, "query x reference block larger than %llu pairs", (unsigned long long)PAIR_CHUNK);
        } else {
            src.first = 0;
            src.a = (const uint32_t *)c->ids_a.p + done;
            src.b = (const uint32_t *)c->ids_b.p + done;
        }
        if ((rc = ensure(c, c->counts, cnt * 4))) return rc;
        CK(cudaMemsetAsync(c->counts.p, 0, cnt * 4, c->stream));
        if (need_pal) {
            if ((rc = ensure(c, c->pal_counts, cnt * 4))) return rc;
            CK(cudaMemsetAsync(c->pal_counts.p, 0, cnt * 4, c->stream));
        }
        if (inter && (rc = ensure(c, c->d_inter, cnt * 8))) return rc;
        if (dist && (rc = ensure(c, c->d_dist, cnt * 8))) return rc;

        CK(cudaEventRecord(c->ev[4], c->stream));
        if (small_main)
            CK(launch_intersect_small((const SetDesc *)c->d_sets.p, src, 0, (uint32_t *)c->counts.p, c->n_sms, c->stream));
        else
            CK(launch_intersect((const SetDesc *)c->d_sets.p, src, 0, (uint32_t)seg, max_segs, (uint32_t *)c->counts.p,
                                (unsigned long long *)c->work_counter.p, c->n_sms, algo, c->stream));
        CK(cudaEventRecord(c->ev[5], c->stream));
        c->m.launches++;
        c->m.intersect_launches++;
        if (need_pal) {
            if (small_pal)
                CK(launch_intersect_small((const SetDesc *)c->d_sets.p, src, 1, (uint32_t *)c->pal_counts.p, c->n_sms,
                                          c->stream));
            else
                CK(launch_intersect((const SetDesc *)c->d_sets.p, src, 1, (uint32_t)seg, max_segs,
                                    (uint32_t *)c->pal_counts.p, (unsigned long long *)c->work_counter.p, c->n_sms,
                                    algo, c->stream));
            c->m.launches++;
        }
        CK(cudaEventRecord(c->ev[6], c->stream));
        CK(launch_epilogue((const SetDesc *)c->d_sets.p, src, (const uint32_t *)c->counts.p,
                           need_pal ? (const uint32_t *)c->pal_counts.p : nullptr, both ? 1 : 0,
                           inter ? (uint64_t *)c->d_inter.p : nullptr, dist ? (double *)c->d_dist.p : nullptr, c->stream));
        c->m.launches++;
        CK(cudaEventRecord(c->ev[7], c->stream));
        if (inter) CK(cudaMemcpyAsync(inter + done, c->d_inter.p, cnt * 8, cudaMemcpyDefault, c->stream));
        if (dist) CK(cudaMemcpyAsync(dist + done, c->d_dist.p, cnt * 8, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->m.d2h_bytes += (inter ? cnt * 8 : 0) + (dist ? cnt * 8 : 0);
        c->m.intersect_ms += elapsed(c->ev[4], c->ev[5]);
        c->m.epilogue_ms += elapsed(c->ev[6], c->ev[7]);
    }
    return GKD_OK;
}

int default_k(int alphabet) { return alphabet == GKD_PROT ? 8 : 21; }

}  // namespace

// =============================================================================================================
// extern "C" ABI
// =============================================================================================================
extern "C" {

int gkd_abi_version(void) { return GKD_ABI_VERSION; }

const char *gkd_last_error(const gkd_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gkd_create(gkd_ctx **out, const gkd_config *cfg) {
    if (!out || !cfg) return fail(nullptr, GKD_EINVAL, "gkd_create: null argument");
    *out = nullptr;
    if (cfg->alphabet < GKD_DNA || cfg->alphabet > GKD_RNA) return fail(nullptr, GKD_EINVAL, "unknown alphabet %d", cfg->alphabet);
    int k = cfg->k == 0 ? default_k(cfg->alphabet) : cfg->k;
    int kmax = cfg->alphabet == GKD_PROT ? 8 : 32;
    if (k < 1 || k > kmax)
        return fail(nullptr, GKD_EINVAL, "kmer size %d is outside the exact-key range 1..%d of this alphabet", k, kmax);
    if (cfg->strand_mode != GKD_STRAND_BOTH && cfg->strand_mode != GKD_STRAND_CANONICAL)
        return fail(nullptr, GKD_EINVAL, "unknown strand mode %d", cfg->strand_mode);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, GKD_ECUDA, "no CUDA device available (%s); libgkd has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, GKD_EINVAL, "device %d out of range (have %d)", cfg->device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess)
        return fail(nullptr, GKD_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, GKD_ECUDA, "device %d is sm_%d%d; libgkd is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
    gkd_ctx *c = new (std::nothrow) gkd_ctx();
    if (!c) return fail(nullptr, GKD_ENOMEM, "out of host memory");
    c->cfg = *cfg;
    c->k = k;
    c->cfg.k = k;
    if (c->cfg.workspace_bytes == 0) c->cfg.workspace_bytes = DEFAULT_WORKSPACE;
    c->n_sms = prop.multiProcessorCount;
#define CK_CREATE(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            fail(nullptr, GKD_ECUDA, "%s: %s", #call, cudaGetErrorString(e__));                \
            delete c;                                                                          \
            return GKD_ECUDA;                                                                  \
        }                                                                                      \
    } while (0)
    CK_CREATE(cudaSetDevice(cfg->device));
    CK_CREATE(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    cudaMemPool_t pool;
    CK_CREATE(cudaDeviceGetDefaultMemPool(&pool, cfg->device));
    uint64_t thresh = UINT64_MAX;
    CK_CREATE(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    for (int i = 0; i < N_STAGE; i++) {
        CK_CREATE(cudaMallocHost((void **)&c->bounce[i], STAGE_BYTES));
        CK_CREATE(cudaEventCreateWithFlags(&c->bounce_ev[i], cudaEventDisableTiming));
        CK_CREATE(cudaMalloc((void **)&c->stage_dev[i], STAGE_BYTES + 256));
    }
    for (auto &ev : c->ev) CK_CREATE(cudaEventCreate(&ev));
    CK_CREATE(intersect_configure());
#undef CK_CREATE
    *out = c;
    return GKD_OK;
}

int gkd_reset(gkd_ctx *c) {
    CHECK_CTX(c);
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    for (auto &a : c->set_arenas) c->free_arenas.push_back(a);
    c->set_arenas.clear();
    c->arena_first_id.clear();
    // keep at most a handful of spare arenas; release the smallest ones beyond that
    while (c->free_arenas.size() > 16) {
        size_t small = 0;
        for (size_t i = 1; i < c->free_arenas.size(); i++)
            if (c->free_arenas[i].second < c->free_arenas[small].second) small = i;
        CK(cudaFreeAsync(c->free_arenas[small].first, c->stream));
        c->free_arenas.erase(c->free_arenas.begin() + small);
    }
    for (auto &s : c->slabs) s.used = 0;
    c->genomes.clear();
    c->built_upto = 0;
    c->sets_dirty = true;
    uint64_t launches = c->m.launches, il = c->m.intersect_launches;
    c->m = gkd_metrics{};
    c->m.launches = launches;
    c->m.intersect_launches = il;
    return GKD_OK;
}

int gkd_truncate(gkd_ctx *c, uint32_t n_keep) {
    CHECK_CTX(c);
    if (n_keep >= c->genomes.size()) return GKD_OK;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    // arenas are created in id order, one per build/import batch: recycle those that hold only dropped sets
    while (!c->set_arenas.empty() && c->arena_first_id.back() >= n_keep) {
        c->free_arenas.push_back(c->set_arenas.back());
        c->set_arenas.pop_back();
        c->arena_first_id.pop_back();
    }
    c->genomes.resize(n_keep);
    if (c->built_upto > n_keep) c->built_upto = n_keep;
    c->sets_dirty = true;
    return GKD_OK;
}

int gkd_destroy(gkd_ctx *c) {
    if (!c) return GKD_EINVAL;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    for (auto &a : c->set_arenas) cudaFreeAsync(a.first, c->stream);
    for (auto &a : c->free_arenas) cudaFreeAsync(a.first, c->stream);
    for (auto &s : c->slabs) cudaFreeAsync(s.base, c->stream);
    DevBuf *bufs[] = {&c->keys_a, &c->keys_b, &c->tile_hist, &c->tile_uniq, &c->genome_counts, &c->batch_genomes,
                      &c->uniq_dst, &c->d_sets, &c->counts, &c->pal_counts, &c->d_inter, &c->d_dist, &c->ids_a,
                      &c->ids_b, &c->work_counter};
    for (DevBuf *b : bufs)
        if (b->p) cudaFreeAsync(b->p, c->stream);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < N_STAGE; i++) {
        if (c->bounce[i]) cudaFreeHost(c->bounce[i]);
        if (c->bounce_ev[i]) cudaEventDestroy(c->bounce_ev[i]);
        if (c->stage_dev[i]) cudaFree(c->stage_dev[i]);
    }
    for (auto &ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
    return GKD_OK;
}

int gkd_add_sequences(gkd_ctx *c, const char *const *contigs, const uint64_t *lens, uint32_t n_contigs, uint32_t *out_id) {
    CHECK_CTX(c);
    if (n_contigs && (!contigs || !lens)) return fail(c, GKD_EINVAL, "gkd_add_sequences: null contig array");
    CK(cudaSetDevice(c->cfg.device));
    std::vector<Part> parts;
    parts.reserve(n_contigs);
    for (uint32_t i = 0; i < n_contigs; i++) {
        if (lens[i] && !contigs[i]) return fail(c, GKD_EINVAL, "gkd_add_sequences: contig %u is null", i);
        parts.push_back(Part{contigs[i], lens[i], true});
    }
    return add_genome(c, parts, "", "", out_id);
}

int gkd_add_fasta_file(gkd_ctx *c, const char *path, int per_record, uint32_t *first_id, uint32_t *n_added) {
    CHECK_CTX(c);
    if (!path) return fail(c, GKD_EINVAL, "gkd_add_fasta_file: null path");
    CK(cudaSetDevice(c->cfg.device));
    std::vector<char> storage;
    std::vector<FastaRecord> recs;
    std::string err;
    if (gkd_parse_fasta_file(path, storage, recs, err)) return fail(c, GKD_EIO, "%s", err.c_str());
    uint32_t first = (uint32_t)c->genomes.size(), added = 0;
    if (per_record) {
        for (auto &r : recs) {
            std::vector<Part> parts;
            bool firstp = true;
            for (auto &l : r.lines) {
                parts.push_back(Part{l.ptr, l.len, firstp});
                firstp = false;
            }
            int rc = add_genome(c, parts, r.label, r.comment, nullptr);
            if (rc) return rc;
            added++;
        }
    } else {
        std::vector<Part> parts;
        for (auto &r : recs) {
            bool firstp = true;
            for (auto &l : r.lines) {
                parts.push_back(Part{l.ptr, l.len, firstp});
                firstp = false;
            }
            if (r.lines.empty()) parts.push_back(Part{"", 0, true});
        }
        std::string label = recs.empty() ? "" : recs[0].label, comment = recs.empty() ? "" : recs[0].comment;
        int rc = add_genome(c, parts, label, comment, nullptr);
        if (rc) return rc;
        added = 1;
    }
    // the staged pieces reference `storage`; they were consumed by memcpy before add_genome returned
    if (first_id) *first_id = first;
    if (n_added) *n_added = added;
    return GKD_OK;
}

const char *gkd_label(const gkd_ctx *c, uint32_t id) { return (c && id < c->genomes.size()) ? c->genomes[id].label.c_str() : ""; }
const char *gkd_comment(const gkd_ctx *c, uint32_t id) { return (c && id < c->genomes.size()) ? c->genomes[id].comment.c_str() : ""; }
uint32_t gkd_count(const gkd_ctx *c) { return c ? (uint32_t)c->genomes.size() : 0; }

int gkd_build_sets(gkd_ctx *c) {
    CHECK_CTX(c);
    CK(cudaSetDevice(c->cfg.device));
    const uint32_t n = (uint32_t)c->genomes.size();
    // skip sets that were imported ready-made
    while (c->built_upto < n) {
        if (c->genomes[c->built_upto].built) {
            c->built_upto++;
            continue;
        }
        // greedy batch under the workspace cap (two key buffers of 8 B per slot)
        const uint64_t cap_keys = std::max<uint64_t>(c->cfg.workspace_bytes / 16, SORT_TILE);
        uint64_t raw = 0;
        uint32_t last = c->built_upto;
        while (last < n && !c->genomes[last].built) {
            uint64_t slots = c->genomes[last].n_pos >= (uint64_t)c->k ? c->genomes[last].n_pos - c->k + 1 : 0;
            slots = (slots + 15) & ~15ull;
            if (last > c->built_upto && raw + slots > cap_keys) break;
            raw += slots;
            last++;
        }
        int rc = build_batch(c, c->built_upto, last);
        if (rc) return rc;
        c->built_upto = last;
    }
    return upload_sets(c);
}

int gkd_set_size(const gkd_ctx *c, uint32_t id, uint64_t *n_both, uint64_t *n_canonical, uint64_t *n_palindromic) {
    if (!c) return GKD_EINVAL;
    if (id >= c->genomes.size() || !c->genomes[id].built) return GKD_ESTATE;
    const SetDesc &d = c->genomes[id].desc;
    const bool both = c->cfg.alphabet != GKD_PROT && c->cfg.strand_mode == GKD_STRAND_BOTH;
    if (n_both) *n_both = both ? 2ull * d.n - d.n_pal : d.n;
    if (n_canonical) *n_canonical = d.n;
    if (n_palindromic) *n_palindromic = d.n_pal;
    return GKD_OK;
}

int gkd_export_set(gkd_ctx *c, uint32_t id, uint64_t *keys, uint64_t cap, uint64_t *n) {
    CHECK_CTX(c);
    int rc = check_built(c, id);
    if (rc) return rc;
    const SetDesc &d = c->genomes[id].desc;
    if (n) *n = d.n;
    if (keys) {
        if (cap < d.n) return fail(c, GKD_EINVAL, "export buffer holds %llu keys, set has %u", (unsigned long long)cap, d.n);
        CK(cudaSetDevice(c->cfg.device));
        if (d.n) CK(cudaMemcpyAsync(keys, d.keys, (uint64_t)d.n * 8, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->m.d2h_bytes += (uint64_t)d.n * 8;
    }
    return GKD_OK;
}

int gkd_set_device_ptr(const gkd_ctx *c, uint32_t id, const uint64_t **keys, uint64_t *n) {
    if (!c) return GKD_EINVAL;
    if (id >= c->genomes.size() || !c->genomes[id].built) return GKD_ESTATE;
    if (keys) *keys = c->genomes[id].desc.keys;
    if (n) *n = c->genomes[id].desc.n;
    return GKD_OK;
}

int gkd_import_sets(gkd_ctx *c, const uint64_t *keys, const uint64_t *offsets, uint32_t n_sets, uint32_t *first_id) {
    CHECK_CTX(c);
    if (n_sets == 0) {
        if (first_id) *first_id = (uint32_t)c->genomes.size();
        return GKD_OK;
    }
    if (!offsets) return fail(c, GKD_EINVAL, "gkd_import_sets: null offsets");
    const uint64_t total = offsets[n_sets];
    if (total && !keys) return fail(c, GKD_EINVAL, "gkd_import_sets: null keys");
    CK(cudaSetDevice(c->cfg.device));
    // treat the arrays as an already-sorted batch and run the unique/compact pass: that re-derives
    // the palindrome lists and drops any duplicate the caller left in
    const uint32_t id0 = (uint32_t)c->genomes.size();
    std::vector<BatchGenome> bg(n_sets);
    std::vector<uint32_t> ids(n_sets);
    uint32_t tiles = 0;
    for (uint32_t i = 0; i < n_sets; i++) {
        if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] >= 0xFFFF0000ull)
            return fail(c, GKD_EINVAL, "gkd_import_sets: bad offsets at set %u", i);
        bg[i].codes = nullptr;
        bg[i].mask = nullptr;
        bg[i].raw_off = offsets[i];
        bg[i].n_pos = 0;
        bg[i].n_slots = (uint32_t)(offsets[i + 1] - offsets[i]);
        bg[i].tile_first = tiles;
        bg[i].n_tiles = (bg[i].n_slots + SORT_TILE - 1) / SORT_TILE;
        tiles += bg[i].n_tiles;
        ids[i] = id0 + i;
    }
    const bool on_device = total && classify(keys) == MEM_DEVICE;
    SortPlan plan{};
    int rc = plan_batch(c, bg, plan, on_device ? 16 : std::max<uint64_t>(total, 16), tiles);
    if (rc) return rc;
    const uint64_t *sorted = keys;
    if (!on_device) {
        if (total) CK(cudaMemcpyAsync(plan.keys_a, keys, total * 8, cudaMemcpyHostToDevice, c->stream));
        sorted = plan.keys_a;
        c->m.h2d_bytes += total * 8;
    }
    for (uint32_t i = 0; i < n_sets; i++) c->genomes.push_back(GenomeRec());
    rc = finish_batch(c, bg, ids, plan, sorted);
    if (rc) {
        c->genomes.resize(id0);
        return rc;
    }
    if (first_id) *first_id = id0;
    return GKD_OK;
}

int gkd_import_set(gkd_ctx *c, const uint64_t *keys, uint64_t n, uint32_t *out_id) {
    uint64_t offsets[2] = {0, n};
    return gkd_import_sets(c, keys, offsets, 1, out_id);
}

// .kset layout (little-endian): "GKDKSET1", u32 k, u32 alphabet, u32 n_sets, u32 0; then per set:
// u64 n_keys, u32 label_len, u32 comment_len, label bytes, comment bytes, n_keys x u64 sorted keys.
int gkd_save_sets(gkd_ctx *c, const char *path) {
    CHECK_CTX(c);
    if (!path) return fail(c, GKD_EINVAL, "gkd_save_sets: null path");
    CK(cudaSetDevice(c->cfg.device));
    for (uint32_t i = 0; i < c->genomes.size(); i++) {
        int rc = check_built(c, i);
        if (rc) return rc;
    }
    FILE *f = fopen(path, "wb");
    if (!f) return fail(c, GKD_EIO, "Cannot open %s for writing.", path);
    uint32_t hdr[4] = {(uint32_t)c->k, (uint32_t)c->cfg.alphabet, (uint32_t)c->genomes.size(), 0};
    bool ok = fwrite("GKDKSET1", 1, 8, f) == 8 && fwrite(hdr, 4, 4, f) == 4;
    std::vector<uint64_t> host;
    for (uint32_t i = 0; ok && i < c->genomes.size(); i++) {
        const GenomeRec &g = c->genomes[i];
        uint64_t n = g.desc.n;
        uint32_t ll[2] = {(uint32_t)g.label.size(), (uint32_t)g.comment.size()};
        host.resize(n);
        if (n) {
            cudaError_t e = cudaMemcpyAsync(host.data(), g.desc.keys, n * 8, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) {
                fclose(f);
                c->poisoned = true;
                return fail(c, GKD_ECUDA, "gkd_save_sets: %s", cudaGetErrorString(e));
            }
        }
        ok = fwrite(&n, 8, 1, f) == 1 && fwrite(ll, 4, 2, f) == 2 &&
             fwrite(g.label.data(), 1, ll[0], f) == ll[0] && fwrite(g.comment.data(), 1, ll[1], f) == ll[1] &&
             fwrite(host.data(), 8, n, f) == n;
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? GKD_OK : fail(c, GKD_EIO, "Write error on %s.", path);
}

int gkd_load_sets(gkd_ctx *c, const char *path, uint32_t *first_id, uint32_t *n_loaded) {
    CHECK_CTX(c);
    if (!path) return fail(c, GKD_EINVAL, "gkd_load_sets: null path");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(c, GKD_EIO, "Input file %s is not found or unreadable.", path);
    char magic[8];
    uint32_t hdr[4];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "GKDKSET1", 8) != 0 || fread(hdr, 4, 4, f) != 4) {
        fclose(f);
        return fail(c, GKD_EIO, "%s is not a .kset file.", path);
    }
    if ((int)hdr[0] != c->k || (int)hdr[1] != c->cfg.alphabet) {
        fclose(f);
        return fail(c, GKD_EINVAL, "%s holds k=%u alphabet=%u sets; this context is k=%d alphabet=%d", path, hdr[0], hdr[1],
                    c->k, c->cfg.alphabet);
    }
    const uint32_t n = hdr[2];
    std::vector<uint64_t> keys, offsets(1, 0);
    std::vector<std::string> labels(n), comments(n);
    bool ok = true;
    for (uint32_t i = 0; ok && i < n; i++) {
        uint64_t nk;
        uint32_t ll[2];
        ok = fread(&nk, 8, 1, f) == 1 && fread(ll, 4, 2, f) == 2 && nk < 0xFFFF0000ull && ll[0] < (1u << 20) && ll[1] < (1u << 20);
        if (!ok) break;
        labels[i].resize(ll[0]);
        comments[i].resize(ll[1]);
        ok = fread(&labels[i][0], 1, ll[0], f) == ll[0] && fread(&comments[i][0], 1, ll[1], f) == ll[1];
        if (!ok) break;
        size_t at = keys.size();
        keys.resize(at + nk);
        ok = fread(keys.data() + at, 8, nk, f) == nk;
        offsets.push_back(keys.size());
    }
    fclose(f);
    if (!ok) return fail(c, GKD_EIO, "%s is truncated or corrupt.", path);
    uint32_t first = 0;
    int rc = gkd_import_sets(c, keys.data(), offsets.data(), n, &first);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; i++) {
        c->genomes[first + i].label = labels[i];
        c->genomes[first + i].comment = comments[i];
    }
    if (first_id) *first_id = first;
    if (n_loaded) *n_loaded = n;
    return GKD_OK;
}

int gkd_all_vs_all(gkd_ctx *c, uint64_t *inter, double *dist) {
    CHECK_CTX(c);
    CK(cudaSetDevice(c->cfg.device));
    uint32_t n = (uint32_t)c->genomes.size();
    HostPairs hp{PAIRS_UPPER, n, 0, n < 2 ? 0 : (uint64_t)n * (n - 1) / 2, nullptr, nullptr, 0, 0};
    return run_pairs(c, hp, inter, dist);
}

int gkd_all_vs_all_range(gkd_ctx *c, uint32_t n, uint64_t first, uint64_t count, uint64_t *inter, double *dist) {
    CHECK_CTX(c);
    CK(cudaSetDevice(c->cfg.device));
    if (n > c->genomes.size()) return fail(c, GKD_EINVAL, "range over %u sets but only %zu exist", n, c->genomes.size());
    uint64_t total = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
    if (first > total || count > total - first) return fail(c, GKD_EINVAL, "pair range [%llu, +%llu) outside 0..%llu", (unsigned long long)first, (unsigned long long)count, (unsigned long long)total);
    HostPairs hp{PAIRS_UPPER, n, first, count, nullptr, nullptr, 0, 0};
    return run_pairs(c, hp, inter, dist);
}

int gkd_query_vs_ref(gkd_ctx *c, const uint32_t *q, uint32_t nq, const uint32_t *r, uint32_t nr, uint64_t *inter, double *dist) {
    CHECK_CTX(c);
    CK(cudaSetDevice(c->cfg.device));
    if ((nq && !q) || (nr && !r)) return fail(c, GKD_EINVAL, "gkd_query_vs_ref: null id array");
    // large blocks are split by query rows so each launch stays under PAIR_CHUNK pairs
    if (nq == 0 || nr == 0) {
        c->m.pairs = 0;
        return GKD_OK;
    }
    uint32_t rows_per = (uint32_t)std::max<uint64_t>(1, PAIR_CHUNK / nr);
    uint64_t pairs = 0, bytes = 0;
    double ims = 0, ems = 0;
    for (uint32_t q0 = 0; q0 < nq; q0 += rows_per) {
        uint32_t rows = std::min(rows_per, nq - q0);
        HostPairs hp{PAIRS_RECT, nr, 0, (uint64_t)rows * nr, q + q0, r, rows, nr};
        int rc = run_pairs(c, hp, inter ? inter + (uint64_t)q0 * nr : nullptr, dist ? dist + (uint64_t)q0 * nr : nullptr);
        if (rc) return rc;
        pairs += c->m.pairs;
        bytes += c->m.intersect_bytes;
        ims += c->m.intersect_ms;
        ems += c->m.epilogue_ms;
    }
    c->m.pairs = pairs;
    c->m.intersect_bytes = bytes;
    c->m.intersect_ms = ims;
    c->m.epilogue_ms = ems;
    return GKD_OK;
}

int gkd_pairs(gkd_ctx *c, const uint32_t *a, const uint32_t *b, uint64_t n_pairs, uint64_t *inter, double *dist) {
    CHECK_CTX(c);
    CK(cudaSetDevice(c->cfg.device));
    if (n_pairs && (!a || !b)) return fail(c, GKD_EINVAL, "gkd_pairs: null id array");
    if (n_pairs > 0xFFFFFFFFull) return fail(c, GKD_EINVAL, "gkd_pairs: more than 2^32 pairs in one call");
    HostPairs hp{PAIRS_LIST, 0, 0, n_pairs, a, b, (uint32_t)n_pairs, (uint32_t)n_pairs};
    return run_pairs(c, hp, inter, dist);
}

int gkd_pair(gkd_ctx *c, uint32_t a, uint32_t b, uint64_t *inter, uint64_t *uni, double *dist) {
    uint64_t I = 0;
    double d = 1.0;
    int rc = gkd_pairs(c, &a, &b, 1, &I, &d);
    if (rc) return rc;
    if (inter) *inter = I;
    if (dist) *dist = d;
    if (uni) {
        uint64_t sa = 0, sb = 0;
        gkd_set_size(c, a, &sa, nullptr, nullptr);
        gkd_set_size(c, b, &sb, nullptr, nullptr);
        *uni = sa + sb - I;
    }
    return GKD_OK;
}

int gkd_get_metrics(const gkd_ctx *c, gkd_metrics *out) {
    if (!c || !out) return GKD_EINVAL;
    *out = c->m;
    return GKD_OK;
}

void *gkd_stream(const gkd_ctx *c) { return c ? (void *)c->stream : nullptr; }

// java.lang.Double.toString: shortest digits that round-trip, decimal layout for 1e-3 <= |v| < 1e7,
// otherwise d.dddE[-]n; always at least one digit after the point.
int gkd_format_double(double v, char *buf, size_t cap) {
    if (std::isnan(v)) return snprintf(buf, cap, "NaN");
    if (std::isinf(v)) return snprintf(buf, cap, v > 0 ? "Infinity" : "-Infinity");
    if (v == 0.0) return snprintf(buf, cap, std::signbit(v) ? "-0.0" : "0.0");
    char tmp[48];
    int prec = 17;
    for (int p = 1; p <= 17; p++) {
        snprintf(tmp, sizeof(tmp), "%.*e", p - 1, v);
        if (strtod(tmp, nullptr) == v) {
            prec = p;
            break;
        }
    }
    snprintf(tmp, sizeof(tmp), "%.*e", prec - 1, v);
    std::string mant;
    const char *e = strchr(tmp, 'e');
    for (const char *p = tmp; p < e; p++)
        if (*p >= '0' && *p <= '9') mant.push_back(*p);
    int ex = atoi(e + 1);
    while (mant.size() > 1 && mant.back() == '0') mant.pop_back();
    std::string out;
    if (v < 0) out.push_back('-');
    if (ex >= -3 && ex < 7) {
        if (ex >= 0) {
            std::string ip = mant.substr(0, std::min<size_t>(mant.size(), (size_t)ex + 1));
            while ((int)ip.size() < ex + 1) ip.push_back('0');
            std::string fp = mant.size() > (size_t)ex + 1 ? mant.substr((size_t)ex + 1) : "0";
            out += ip + "." + fp;
        } else {
            out += "0." + std::string((size_t)(-ex - 1), '0') + mant;
        }
    } else {
        out += mant.substr(0, 1) + "." + (mant.size() > 1 ? mant.substr(1) : std::string("0")) + "E" + std::to_string(ex);
    }
    return snprintf(buf, cap, "%s", out.c_str());
}

static int synth_any(int device, char *dst, uint64_t len, uint64_t seed, uint32_t family, uint32_t member, double rate,
                     int protein) {
    if (len && !dst) return GKD_EINVAL;
    if (device >= 0 && classify(dst) == MEM_DEVICE) {
        if (cudaSetDevice(device) != cudaSuccess) return GKD_ECUDA;
        if (launch_synth(dst, len, seed, family, member, rate, protein, 0) != cudaSuccess) return GKD_ECUDA;
        if (cudaStreamSynchronize(0) != cudaSuccess) return GKD_ECUDA;
        return GKD_OK;
    }
    synth_host(dst, len, seed, family, member, rate, protein);
    return GKD_OK;
}

int gkd_synth_dna(int device, char *dst, uint64_t len, uint64_t seed, uint32_t family, uint32_t member, double sub_rate) {
    return synth_any(device, dst, len, seed, family, member, sub_rate, 0);
}

int gkd_synth_protein(int device, char *dst, uint64_t len, uint64_t seed, uint32_t family, uint32_t member, double sub_rate) {
    return synth_any(device, dst, len, seed, family, member, sub_rate, 1);
}

}  // extern "C"
