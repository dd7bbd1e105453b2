// kmers.hpp -- C++ host-side mirror of the reference's k-mer object API over the gkd C ABI.
//
// The reference is Java and the image has no JVM, so the host layer above the C ABI is written in
// C++ with the reference's names, argument meaning and error behaviour (java/ holds the JNI source
// a maintainer would compile where a JDK exists; INTEGRATION.md shows the binding).
//
//   reference (org.theseed.sequence, external)            here
//   -------------------------------------------------     -----------------------------------------
//   KmerType.DNA / getKmerSize() / createKmers(seq, K)     KmerType, kmerTypeDefaultK, KmerEngine::createKmers
//   SequenceKmers.distance(other)                          SequenceKmers::distance
//   new GenomeKmers(genome), getGenomeId/Name              KmerEngine::genomeKmers, SequenceKmers::id/name
//   new ProteinKmers(str)                                  KmerEngine::createKmers with KmerType::PROT
//   GenomeKmers.setKmerSize / ProteinKmers.setKmerSize      per-engine K (the static global becomes ctx state)
//
// Per-pair calls stay available (SequenceKmers::distance -> gkd_pair) for the greedy callers
// (DistanceRepsProcessor.java:101,190; FastaDistanceRepsProcessor.java:128), but the processors use
// the batched entry points: one JNI crossing + one launch per pair would be launch-bound.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "gkd.h"

namespace theseed {

// org.theseed.basic.ParseFailureException
struct ParseFailureException : std::runtime_error {
    using std::runtime_error::runtime_error;
};
// java.io.FileNotFoundException / IOException
struct IOException : std::runtime_error {
    using std::runtime_error::runtime_error;
};

enum class KmerType { DNA = GKD_DNA, PROT = GKD_PROT, RNA = GKD_RNA };

// KmerType.getKmerSize(): 21 for DNA/RNA, 8 for protein (FastaDistanceProcessor.java:43)
inline int kmerTypeDefaultK(KmerType t) { return t == KmerType::PROT ? 8 : 21; }

inline KmerType parseKmerType(const std::string &s) {
    if (s == "DNA") return KmerType::DNA;
    if (s == "RNA") return KmerType::RNA;
    if (s == "PROT" || s == "PROTEIN") return KmerType::PROT;
    throw ParseFailureException("\"" + s + "\" is not a valid value for \"--type\"");
}

class KmerEngine;

// handle to one device-resident k-mer set
class SequenceKmers {
  public:
    SequenceKmers() = default;
    SequenceKmers(KmerEngine *e, uint32_t id, std::string gid = "", std::string name = "")
        : eng_(e), id_(id), gid_(std::move(gid)), name_(std::move(name)) {}
    double distance(const SequenceKmers &other) const;  // SequenceKmers.distance
    uint64_t similarity(const SequenceKmers &other) const;
    uint64_t size() const;
    uint32_t handle() const { return id_; }
    const std::string &getGenomeId() const { return gid_; }
    const std::string &getGenomeName() const { return name_; }

  private:
    KmerEngine *eng_ = nullptr;
    uint32_t id_ = 0;
    std::string gid_, name_;
};

class KmerEngine {
  public:
    KmerEngine(KmerType type, int k, int device = 0) : type_(type) {
        gkd_config cfg{};
        cfg.device = device;
        cfg.k = k;
        cfg.alphabet = (int)type;
        cfg.strand_mode = GKD_STRAND_BOTH;
        int rc = gkd_create(&ctx_, &cfg);
        if (rc) raise(rc, gkd_last_error(nullptr));
    }
    ~KmerEngine() {
        if (ctx_) gkd_destroy(ctx_);
    }
    KmerEngine(const KmerEngine &) = delete;
    KmerEngine &operator=(const KmerEngine &) = delete;

    // KmerType.createKmers(seq, K) / new ProteinKmers(str): queued; sets materialise at build()
    SequenceKmers createKmers(const std::string &seq) {
        const char *p = seq.data();
        uint64_t n = seq.size();
        uint32_t id = 0;
        check(gkd_add_sequences(ctx_, &p, &n, 1, &id));
        return SequenceKmers(this, id);
    }
    // new GenomeKmers(genome): one piece per contig
    SequenceKmers genomeKmers(const std::vector<std::string> &contigs, const std::string &gid, const std::string &name) {
        std::vector<const char *> ptrs;
        std::vector<uint64_t> lens;
        for (auto &c : contigs) {
            ptrs.push_back(c.data());
            lens.push_back(c.size());
        }
        uint32_t id = 0;
        check(gkd_add_sequences(ctx_, ptrs.data(), lens.data(), (uint32_t)ptrs.size(), &id));
        return SequenceKmers(this, id, gid, name);
    }
    // FastaInputStream: every record of a FASTA file as its own sequence
    std::vector<SequenceKmers> addFasta(const std::string &path) {
        uint32_t first = 0, n = 0;
        check(gkd_add_fasta_file(ctx_, path.c_str(), 1, &first, &n));
        std::vector<SequenceKmers> out;
        for (uint32_t i = 0; i < n; i++) out.emplace_back(this, first + i, gkd_label(ctx_, first + i), gkd_comment(ctx_, first + i));
        return out;
    }
    void build() { check(gkd_build_sets(ctx_)); }
    uint32_t count() const { return gkd_count(ctx_); }
    void allVsAll(std::vector<uint64_t> *inter, std::vector<double> &dist) {
        uint64_t n = count(), np = n < 2 ? 0 : n * (n - 1) / 2;
        dist.assign(np, 1.0);
        if (inter) inter->assign(np, 0);
        check(gkd_all_vs_all(ctx_, inter ? inter->data() : nullptr, dist.data()));
    }
    void queryVsRef(const std::vector<uint32_t> &q, const std::vector<uint32_t> &r, std::vector<double> &dist) {
        dist.assign((uint64_t)q.size() * r.size(), 1.0);
        check(gkd_query_vs_ref(ctx_, q.data(), (uint32_t)q.size(), r.data(), (uint32_t)r.size(), nullptr, dist.data()));
    }
    gkd_ctx *raw() { return ctx_; }
    KmerType type() const { return type_; }
    void check(int rc) {
        if (rc) raise(rc, gkd_last_error(ctx_));
    }

  private:
    [[noreturn]] static void raise(int rc, const char *msg) {
        std::string m = msg ? msg : "unknown error";
        if (rc == GKD_EINVAL) throw ParseFailureException(m);  // mirrors GenomeProcessor.java:112-115
        if (rc == GKD_EIO) throw IOException(m);
        throw std::runtime_error(m);
    }
    gkd_ctx *ctx_ = nullptr;
    KmerType type_;
};

inline double SequenceKmers::distance(const SequenceKmers &o) const {
    double d = 1.0;
    eng_->check(gkd_pair(eng_->raw(), id_, o.id_, nullptr, nullptr, &d));
    return d;
}
inline uint64_t SequenceKmers::similarity(const SequenceKmers &o) const {
    uint64_t i = 0;
    eng_->check(gkd_pair(eng_->raw(), id_, o.id_, &i, nullptr, nullptr));
    return i;
}
inline uint64_t SequenceKmers::size() const {
    uint64_t n = 0;
    eng_->check(gkd_set_size(eng_->raw(), id_, &n, nullptr, nullptr));
    return n;
}

}  // namespace theseed
