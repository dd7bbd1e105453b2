// fasta.cpp -- host FASTA scanner of libgkd.so.
// Restates org.theseed.sequence.FastaInputStream as used by FastaDistanceProcessor.java:104-108,119-131:
// a record starts at a '>' line; label = header text up to the first whitespace, comment = the rest
// (trimmed); the sequence is the concatenation of the following lines with line ends removed.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <sys/stat.h>
#include <string>
#include <vector>

struct FastaPiece {
    const char *ptr;
    uint64_t len;
};
struct FastaRecord {
    std::string label, comment;
    std::vector<FastaPiece> lines;
};

static inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; }

int gkd_parse_fasta_file(const char *path, std::vector<char> &storage, std::vector<FastaRecord> &records,
                         std::string &err) {
    FILE *f = (strcmp(path, "-") == 0) ? stdin : fopen(path, "rb");
    if (!f) {
        err = std::string("Input file ") + path + " is not found or unreadable.";
        return -1;
    }
    storage.clear();
    bool bad = false;
    if (f != stdin) {
        // size the buffer once and read the file in one call
        struct stat st;
        if (fstat(fileno(f), &st) == 0 && st.st_size > 0) {
            storage.resize((size_t)st.st_size);
            size_t got = fread(storage.data(), 1, storage.size(), f);
            storage.resize(got);
        }
    }
    {   // stdin, or whatever is left (size unknown / file grew)
        char buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof(buf), f)) > 0) storage.insert(storage.end(), buf, buf + got);
        bad = ferror(f) != 0;
    }
    if (f != stdin) fclose(f);
    if (bad) {
        err = std::string("Read error on ") + path + ".";
        return -1;
    }
    records.clear();
    // One pass; the sequence lines of a record are compacted in place (line ends and surrounding
    // white space dropped) so every record ends up as ONE contiguous piece of `storage`.
    char *base = storage.data();
    char *p = base;
    char *end = base + storage.size();
    char *w = base;          // write cursor of the compaction (never ahead of the read cursor)
    char *seq_start = nullptr;
    auto close_record = [&]() {
        if (!records.empty() && seq_start && w > seq_start)
            records.back().lines.push_back(FastaPiece{seq_start, (uint64_t)(w - seq_start)});
        seq_start = nullptr;
    };
    while (p < end) {
        char *nl = (char *)memchr(p, '\n', (size_t)(end - p));
        char *le = nl ? nl : end;
        char *a = p, *b = le;
        while (a < b && is_space(*a)) a++;
        while (b > a && is_space(b[-1])) b--;
        if (a < b) {
            if (*a == '>') {
                close_record();
                FastaRecord r;
                const char *h = a + 1;
                while (h < b && is_space(*h)) h++;
                const char *ws = h;
                while (ws < b && !is_space(*ws)) ws++;
                r.label.assign(h, ws);
                while (ws < b && is_space(*ws)) ws++;
                r.comment.assign(ws, (const char *)b);
                records.push_back(std::move(r));
                w = (nl ? nl + 1 : end);  // the sequence is compacted right after its header line
                seq_start = w;
            } else if (!records.empty()) {
                size_t len = (size_t)(b - a);
                if (w != a) memmove(w, a, len);
                w += len;
            }
        }
        p = nl ? nl + 1 : end;
    }
    close_record();
    return 0;
}
