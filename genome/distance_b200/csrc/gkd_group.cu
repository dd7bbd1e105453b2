// gkd_group.cu -- several devices behind the C ABI: a group of contexts (one per listed device) in ONE
// process, one host thread per member, set arenas moved with peer copies (copy engines over NVLink) and
// adopted in place.  Built entirely on the public single-device ABI (gkd.h) plus cudaMemcpyPeerAsync.
//
// Reference shape: FastaDistanceProcessor.runReporter / computePairs (:141-194) -- all pairs i < j of one
// FASTA file.  The pair matrix shards with no reduction: member x owns its diagonal
// block and the blocks (x, (x+s) mod R), s = 1..R/2 (the half-way block of an even ring is split between
// its two members), exactly the schedule of genome/distance_b200/sharding.py (the one-process-per-GPU
// torch.distributed layer); here the transport is a peer copy instead of NCCL send/recv.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "gkd.h"

// host FASTA scanner (fasta.cpp)
struct FastaPiece {
    const char *ptr;
    uint64_t len;
};
struct FastaRecord {
    std::string label, comment;
    std::vector<FastaPiece> lines;
};
int gkd_parse_fasta_file(const char *path, std::vector<char> &storage, std::vector<FastaRecord> &records,
                         std::string &err);

namespace {

struct ArenaMeta {  // snapshot of one arena of a member after the build (read-only during the ring)
    uint32_t first, n;
    const char *base;
    uint64_t bytes;
    std::vector<gkd_packed_set> table;  // offsets relative to base
};

struct Panel {  // sets [first, first+count) of a member: bytes [begin, end) of one of its arenas
    uint32_t arena, first, count;
    uint64_t begin, end;
    std::vector<gkd_packed_set> table;  // offsets relative to begin
};

struct Member {
    gkd_ctx *ctx = nullptr;
    int device = 0;
    std::vector<uint32_t> global_ids;  // local set id -> global id (ascending: members own contiguous id blocks)
    std::vector<ArenaMeta> arenas;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    void *recv[2] = {nullptr, nullptr};
    uint64_t recv_cap[2] = {0, 0};
    int rc = GKD_OK;
    std::string err;
};

thread_local std::string g_group_create_error;

}  // namespace

struct gkd_group {
    gkd_config cfg{};
    std::vector<Member> members;
    uint32_t n_genomes = 0;
    uint32_t panel_sets = 128;
    std::vector<std::pair<uint32_t, uint32_t>> where;  // global id -> (member, local id)
    std::string err;
};

namespace {

int gfail(gkd_group *g, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (g) g->err = buf;
    else g_group_create_error = buf;
    return code;
}

// plan_panels of sharding.py: never spans two arenas, at most max_sets sets, offsets rebased to `begin`
std::vector<Panel> plan_panels(const std::vector<ArenaMeta> &meta, uint32_t lo, uint32_t hi, uint32_t max_sets) {
    std::vector<Panel> out;
    for (uint32_t ai = 0; ai < meta.size(); ai++) {
        const ArenaMeta &A = meta[ai];
        uint32_t a = std::max(lo, A.first), b = std::min(hi, A.first + A.n);
        while (a < b) {
            const uint32_t m = std::min(b - a, std::max(1u, max_sets));
            Panel p;
            p.arena = ai;
            p.first = a;
            p.count = m;
            p.table.assign(A.table.begin() + (a - A.first), A.table.begin() + (a - A.first + m));
            p.begin = p.table[0].offs_off;
            p.end = (a + m == A.first + A.n) ? A.bytes : A.table[a - A.first + m].offs_off;
            for (auto &t : p.table) {
                t.offs_off -= p.begin;
                t.lows_off -= p.begin;
                if (t.pal_offs_off) t.pal_offs_off -= p.begin;
                if (t.pal_lows_off) t.pal_lows_off -= p.begin;
            }
            out.push_back(std::move(p));
            a += m;
        }
    }
    return out;
}

inline uint64_t row_start(uint64_t i, uint64_t n) { return i * (2 * n - i - 1) / 2; }

#define MCK(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            char b__[256];                                                                                \
            snprintf(b__, sizeof(b__), "%s: %s", #call, cudaGetErrorString(e__));                         \
            M.err = b__;                                                                                  \
            M.rc = e__ == cudaErrorMemoryAllocation ? GKD_ENOMEM : GKD_ECUDA;                             \
            return;                                                                                       \
        }                                                                                                 \
    } while (0)
#define MGK(call)                                             \
    do {                                                      \
        int r__ = (call);                                     \
        if (r__ != GKD_OK) {                                  \
            M.rc = r__;                                       \
            M.err = gkd_last_error(M.ctx);                    \
            return;                                           \
        }                                                     \
    } while (0)

void member_ring(gkd_group *g, uint32_t x, uint64_t *inter, double *dist);

// the ring of one member (runs on its own host thread; no exception may leave the thread)
void member_all_vs_all(gkd_group *g, uint32_t x, uint64_t *inter, double *dist) {
    try {
        member_ring(g, x, inter, dist);
    } catch (const std::exception &ex) {
        g->members[x].rc = GKD_ENOMEM;
        g->members[x].err = ex.what();
    }
}

void member_ring(gkd_group *g, uint32_t x, uint64_t *inter, double *dist) {
    Member &M = g->members[x];
    const uint32_t R = (uint32_t)g->members.size();
    const uint64_t N = g->n_genomes;
    const uint32_t m = (uint32_t)M.global_ids.size();
    MCK(cudaSetDevice(M.device));
    std::vector<uint64_t> bi;
    std::vector<double> bd;
    auto put = [&](uint32_t ga, uint32_t gb, uint64_t I, double d) {
        if (ga > gb) std::swap(ga, gb);
        const uint64_t t = row_start(ga, N) + (gb - ga - 1);
        if (inter) inter[t] = I;
        if (dist) dist[t] = d;
    };
    // transfer slots: (panel to receive, its owner, my rows)
    struct Slot {
        Panel p;
        uint32_t src, row0, row1;
    };
    std::vector<Slot> slots;
    for (uint32_t s = 1; s <= R / 2; s++) {
        const uint32_t src = (x + s) % R;
        const bool half = (R % 2 == 0) && s == R / 2;
        const uint32_t ms = (uint32_t)g->members[src].global_ids.size();
        uint32_t row0 = 0, row1 = m, rc0 = 0, rc1 = ms;
        if (half) {  // block X x Y (X = lower member): lower computes X[:h] x Y, higher computes Y x X[h:]
            if (x < src) row1 = (m + 1) / 2;
            else rc0 = (ms + 1) / 2;
        }
        if (row1 <= row0) continue;
        for (auto &p : plan_panels(g->members[src].arenas, rc0, rc1, g->panel_sets)) slots.push_back(Slot{std::move(p), src, row0, row1});
    }
    // Panels are the unit of transfer; the unit of compute is a GROUP of consecutive panels that meet the same rows, up
    // to GROUP_SETS sets (possibly of several peers): kernel 4's block join builds one table per (64 rows, key range) and
    // probes every column of the call with it, so one call over many columns amortises the tables.
    constexpr uint32_t GROUP_SETS = 1024;
    struct Group {
        size_t s0, s1;  // slots [s0, s1)
        std::vector<uint64_t> off;  // byte offset of every panel in the receive buffer
        uint64_t bytes;
    };
    std::vector<Group> groups;
    for (size_t k = 0; k < slots.size();) {
        Group G{k, k, {}, 0};
        uint32_t sets = 0;
        while (G.s1 < slots.size() && slots[G.s1].row0 == slots[k].row0 && slots[G.s1].row1 == slots[k].row1 &&
               (G.s1 == k || sets + slots[G.s1].p.count <= GROUP_SETS)) {
            G.off.push_back(G.bytes);
            G.bytes += (slots[G.s1].p.end - slots[G.s1].p.begin + 255) & ~255ull;
            sets += slots[G.s1].p.count;
            G.s1++;
        }
        groups.push_back(std::move(G));
        k = groups.back().s1;
    }
    auto post = [&](size_t gi) {  // start the peer copies of group gi into receive buffer gi & 1
        const Group &G = groups[gi];
        const int b = (int)(gi & 1);
        if (G.bytes > M.recv_cap[b]) {
            if (M.recv[b]) MCK(cudaFree(M.recv[b]));
            M.recv[b] = nullptr;
            M.recv_cap[b] = 0;
            MCK(cudaMalloc(&M.recv[b], G.bytes + G.bytes / 8 + 256));
            M.recv_cap[b] = G.bytes + G.bytes / 8 + 256;
        }
        for (size_t k = G.s0; k < G.s1; k++) {
            const Slot &S = slots[k];
            const Member &O = g->members[S.src];
            const char *src_ptr = O.arenas[S.p.arena].base + S.p.begin;
            MCK(cudaMemcpyPeerAsync((char *)M.recv[b] + G.off[k - G.s0], M.device, src_ptr, O.device, S.p.end - S.p.begin, M.copy_stream));
        }
        MCK(cudaEventRecord(M.copy_done[b], M.copy_stream));
    };
    if (!groups.empty()) post(0);
    if (M.rc) return;

    // diagonal block (runs while the first group is in flight)
    if (m >= 2) {
        const uint64_t cnt = (uint64_t)m * (m - 1) / 2;
        bi.resize(cnt);
        bd.resize(cnt);
        MGK(gkd_all_vs_all_range(M.ctx, m, 0, cnt, bi.data(), bd.data()));
        uint64_t t = 0;
        for (uint32_t i = 0; i < m; i++)
            for (uint32_t j = i + 1; j < m; j++, t++) put(M.global_ids[i], M.global_ids[j], bi[t], bd[t]);
    }
    std::vector<uint32_t> rows, cols;
    for (size_t gi = 0; gi < groups.size(); gi++) {
        if (gi + 1 < groups.size()) {
            // buffer (gi+1)&1 was last used by group gi-1, whose sets were dropped (and the stream drained) below
            post(gi + 1);
            if (M.rc) return;
        }
        const Group &G = groups[gi];
        const int b = (int)(gi & 1);
        MCK(cudaEventSynchronize(M.copy_done[b]));
        const Slot &S0 = slots[G.s0];
        rows.clear();
        cols.clear();
        for (uint32_t r = S0.row0; r < S0.row1; r++) rows.push_back(r);
        for (size_t k = G.s0; k < G.s1; k++) {
            const Slot &S = slots[k];
            uint32_t first = 0;
            MGK(gkd_adopt_sets(M.ctx, (char *)M.recv[b] + G.off[k - G.s0], S.p.end - S.p.begin, S.p.table.data(), S.p.count, &first));
            for (uint32_t c = 0; c < S.p.count; c++) cols.push_back(first + c);
        }
        bi.resize((size_t)rows.size() * cols.size());
        bd.resize(bi.size());
        MGK(gkd_query_vs_ref(M.ctx, rows.data(), (uint32_t)rows.size(), cols.data(), (uint32_t)cols.size(), bi.data(), bd.data()));
        size_t col0 = 0;
        for (size_t k = G.s0; k < G.s1; k++) {
            const Slot &S = slots[k];
            const std::vector<uint32_t> &their = g->members[S.src].global_ids;
            for (uint32_t r = S0.row0; r < S0.row1; r++)
                for (uint32_t c = 0; c < S.p.count; c++) {
                    const size_t t = (size_t)(r - S0.row0) * cols.size() + col0 + c;
                    put(M.global_ids[r], their[S.p.first + c], bi[t], bd[t]);
                }
            col0 += S.p.count;
        }
        MGK(gkd_truncate(M.ctx, m));
    }
}

}  // namespace

extern "C" {

const char *gkd_group_last_error(const gkd_group *g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }

int gkd_group_create(gkd_group **out, const gkd_config *cfg, const int32_t *devices, uint32_t n_devices) {
    if (!out || !cfg || !devices || n_devices == 0) return gfail(nullptr, GKD_EINVAL, "gkd_group_create: null or empty argument");
    *out = nullptr;
    gkd_group *g = new (std::nothrow) gkd_group();
    if (!g) return gfail(nullptr, GKD_ENOMEM, "out of host memory");
    try {
        g->cfg = *cfg;
        g->members.resize(n_devices);
        for (uint32_t i = 0; i < n_devices; i++) {
            Member &M = g->members[i];
            M.device = devices[i];
            gkd_config c = *cfg;
            c.device = devices[i];
            int rc = gkd_create(&M.ctx, &c);
            if (rc != GKD_OK) {
                gfail(nullptr, rc, "member %u (device %d): %s", i, devices[i], gkd_last_error(nullptr));
                gkd_group_destroy(g);
                return rc;
            }
            bool ok = cudaSetDevice(M.device) == cudaSuccess &&
                      cudaStreamCreateWithFlags(&M.copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
                      cudaEventCreateWithFlags(&M.copy_done[0], cudaEventDisableTiming) == cudaSuccess &&
                      cudaEventCreateWithFlags(&M.copy_done[1], cudaEventDisableTiming) == cudaSuccess;
            if (!ok) {
                gfail(nullptr, GKD_ECUDA, "member %u (device %d): cannot create the copy stream", i, devices[i]);
                gkd_group_destroy(g);
                return GKD_ECUDA;
            }
        }
        // direct peer copies over NVLink where the hardware allows it (staged through the host otherwise)
        for (uint32_t i = 0; i < n_devices; i++)
            for (uint32_t j = 0; j < n_devices; j++) {
                if (devices[i] == devices[j]) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
                    cudaSetDevice(devices[i]);
                    cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                    if (e != cudaSuccess) cudaGetLastError();  // already enabled is fine
                }
            }
    } catch (const std::bad_alloc &) {
        gkd_group_destroy(g);
        return gfail(nullptr, GKD_ENOMEM, "out of host memory");
    }
    *out = g;
    return GKD_OK;
}

int gkd_group_destroy(gkd_group *g) {
    if (!g) return GKD_EINVAL;
    for (auto &M : g->members) {
        if (M.ctx) {
            cudaSetDevice(M.device);
            if (M.copy_stream) cudaStreamSynchronize(M.copy_stream);
            gkd_destroy(M.ctx);  // drops the adopted sets before their buffers go
        }
        cudaSetDevice(M.device);
        for (int b = 0; b < 2; b++) {
            if (M.recv[b]) cudaFree(M.recv[b]);
            if (M.copy_done[b]) cudaEventDestroy(M.copy_done[b]);
        }
        if (M.copy_stream) cudaStreamDestroy(M.copy_stream);
    }
    cudaGetLastError();
    delete g;
    return GKD_OK;
}

uint32_t gkd_group_size(const gkd_group *g) { return g ? (uint32_t)g->members.size() : 0; }
uint32_t gkd_group_count(const gkd_group *g) { return g ? g->n_genomes : 0; }
gkd_ctx *gkd_group_member(gkd_group *g, uint32_t member) { return (g && member < g->members.size()) ? g->members[member].ctx : nullptr; }

int gkd_group_set_panel(gkd_group *g, uint32_t sets_per_panel) {
    if (!g || sets_per_panel == 0) return GKD_EINVAL;
    g->panel_sets = sets_per_panel;
    return GKD_OK;
}

int gkd_group_add_sequences(gkd_group *g, uint32_t member, const char *const *contigs, const uint64_t *lens, uint32_t n_contigs,
                            uint32_t *global_id) {
    if (!g) return GKD_EINVAL;
    if (member >= g->members.size()) return gfail(g, GKD_EINVAL, "member %u out of range (have %zu)", member, g->members.size());
    // members own contiguous blocks of global ids: a genome may only go to the current member or a later one
    for (uint32_t later = member + 1; later < g->members.size(); later++)
        if (!g->members[later].global_ids.empty())
            return gfail(g, GKD_EINVAL, "genomes must be added member by member in ascending member order");
    try {
        Member &M = g->members[member];
        uint32_t local = 0;
        int rc = gkd_add_sequences(M.ctx, contigs, lens, n_contigs, &local);
        if (rc != GKD_OK) return gfail(g, rc, "%s", gkd_last_error(M.ctx));
        M.global_ids.push_back(g->n_genomes);
        g->where.push_back({member, local});
        if (global_id) *global_id = g->n_genomes;
        g->n_genomes++;
        return GKD_OK;
    } catch (const std::bad_alloc &) {
        return gfail(g, GKD_ENOMEM, "out of host memory");
    }
}

int gkd_group_add_fasta_file(gkd_group *g, const char *path, uint32_t *n_added) {
    if (!g || !path) return GKD_EINVAL;
    if (g->n_genomes != 0) return gfail(g, GKD_ESTATE, "gkd_group_add_fasta_file needs an empty group (it block-distributes the records)");
    try {
        std::vector<char> storage;
        std::vector<FastaRecord> recs;
        std::string err;
        if (gkd_parse_fasta_file(path, storage, recs, err)) return gfail(g, GKD_EIO, "%s", err.c_str());
        const uint32_t R = (uint32_t)g->members.size(), n = (uint32_t)recs.size();
        const uint32_t base = n / R, extra = n % R;  // block distribution, sizes differ by at most one
        uint32_t r = 0;
        for (uint32_t member = 0; member < R; member++) {
            const uint32_t take = base + (member < extra ? 1 : 0);
            for (uint32_t i = 0; i < take; i++, r++) {
                // one record = one genome whose sequence lines are pieces of ONE contig: hand the pieces through
                // the member context's FASTA-aware path by concatenating them (records are small relative to the file)
                std::string seq;
                for (auto &l : recs[r].lines) seq.append(l.ptr, l.len);
                const char *p = seq.data();
                uint64_t len = seq.size();
                int rc = gkd_group_add_sequences(g, member, &p, &len, 1, nullptr);
                if (rc != GKD_OK) return rc;
                gkd_set_label(g->members[member].ctx, g->where.back().second, recs[r].label.c_str(), recs[r].comment.c_str());
            }
        }
        if (n_added) *n_added = n;
        return GKD_OK;
    } catch (const std::bad_alloc &) {
        return gfail(g, GKD_ENOMEM, "out of host memory");
    }
}

const char *gkd_group_label(const gkd_group *g, uint32_t global_id) {
    if (!g || global_id >= g->where.size()) return "";
    return gkd_label(g->members[g->where[global_id].first].ctx, g->where[global_id].second);
}
const char *gkd_group_comment(const gkd_group *g, uint32_t global_id) {
    if (!g || global_id >= g->where.size()) return "";
    return gkd_comment(g->members[g->where[global_id].first].ctx, g->where[global_id].second);
}

int gkd_group_build(gkd_group *g) {
    if (!g) return GKD_EINVAL;
    try {
        std::vector<std::thread> threads;
        for (auto &M : g->members) {
            M.rc = GKD_OK;
            threads.emplace_back([&M]() {
                M.rc = gkd_build_sets(M.ctx);
                if (M.rc != GKD_OK) M.err = gkd_last_error(M.ctx);
            });
        }
        for (auto &t : threads) t.join();
        for (size_t i = 0; i < g->members.size(); i++)
            if (g->members[i].rc != GKD_OK) return gfail(g, g->members[i].rc, "member %zu: %s", i, g->members[i].err.c_str());
        // snapshot every member's arena layout: the ring reads its peers' layouts while they adopt and drop panels
        for (auto &M : g->members) {
            M.arenas.clear();
            const uint32_t na = gkd_arena_count(M.ctx);
            for (uint32_t a = 0; a < na; a++) {
                ArenaMeta A;
                const void *base = nullptr;
                int rc = gkd_arena_info(M.ctx, a, &A.first, &A.n, &base, &A.bytes);
                if (rc != GKD_OK) return gfail(g, rc, "gkd_arena_info failed");
                A.base = (const char *)base;
                A.table.resize(A.n);
                rc = gkd_describe_sets(M.ctx, A.first, A.n, A.table.data());
                if (rc != GKD_OK) return gfail(g, rc, "%s", gkd_last_error(M.ctx));
                M.arenas.push_back(std::move(A));
            }
        }
        return GKD_OK;
    } catch (const std::exception &ex) {
        return gfail(g, GKD_ENOMEM, "gkd_group_build: %s", ex.what());
    }
}

int gkd_group_all_vs_all(gkd_group *g, uint64_t *inter, double *dist) {
    if (!g) return GKD_EINVAL;
    try {
        for (auto &M : g->members)
            if (M.arenas.empty() && !M.global_ids.empty()) return gfail(g, GKD_ESTATE, "call gkd_group_build first");
        std::vector<std::thread> threads;
        for (uint32_t x = 0; x < g->members.size(); x++) {
            g->members[x].rc = GKD_OK;
            threads.emplace_back(member_all_vs_all, g, x, inter, dist);
        }
        for (auto &t : threads) t.join();
        for (size_t i = 0; i < g->members.size(); i++)
            if (g->members[i].rc != GKD_OK) return gfail(g, g->members[i].rc, "member %zu: %s", i, g->members[i].err.c_str());
        return GKD_OK;
    } catch (const std::exception &ex) {
        return gfail(g, GKD_ENOMEM, "gkd_group_all_vs_all: %s", ex.what());
    }
}

}  // extern "C"
