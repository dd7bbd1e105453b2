// fasta.cpp -- host FASTA scanner of libgkd.so.
// Restates org.theseed.sequence.FastaInputStream as used by FastaDistanceProcessor.java:104-108,119-131:
// a record starts at a '>' line; label = header text up to the first whitespace, comment = the rest
// (trimmed); the sequence is the concatenation of the following lines with line ends removed.
#include <cstdint>
#include <cstdio>
#include <cstring>
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
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) storage.insert(storage.end(), buf, buf + got);
    bool bad = ferror(f) != 0;
    if (f != stdin) fclose(f);
    if (bad) {
        err = std::string("Read error on ") + path + ".";
        return -1;
    }
    records.clear();
    const char *p = storage.data();
    const char *end = p + storage.size();
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        const char *a = p, *b = le;
        while (a < b && is_space(*a)) a++;
        while (b > a && is_space(b[-1])) b--;
        if (a < b) {
            if (*a == '>') {
                FastaRecord r;
                const char *h = a + 1;
                while (h < b && is_space(*h)) h++;
                const char *ws = h;
                while (ws < b && !is_space(*ws)) ws++;
                r.label.assign(h, ws);
                while (ws < b && is_space(*ws)) ws++;
                r.comment.assign(ws, b);
                records.push_back(std::move(r));
            } else if (!records.empty()) {
                records.back().lines.push_back(FastaPiece{a, (uint64_t)(b - a)});
            }
        }
        p = nl ? nl + 1 : end;
    }
    return 0;
}
