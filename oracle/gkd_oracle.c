/*
 * gkd_oracle.c -- CPU ORACLE (test infrastructure only; see gkd_oracle.h for scope and provenance).
 *
 * PARITY UNPINNED: no golden vectors exist in /root/reference; the arithmetic lives in the
 * un-vendored org.theseed:sequence:1.0.0.  Every function cites the reference call site whose
 * contract it restates.
 */
#include "gkd_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------ */
/* character handling                                                                         */
/* ------------------------------------------------------------------------------------------ */

static inline char lower_ascii(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c; }

/* DnaKmers lower-cases its input (String.toLowerCase); RNA additionally reads u as t */
static inline char fold_nuc(char c, int alphabet) {
    c = lower_ascii(c);
    if (alphabet == ORC_RNA && c == 'u') c = 't';
    return c;
}

static inline int nuc_code(char c) {
    switch (c) {
    case 'a': return 0;
    case 'c': return 1;
    case 'g': return 2;
    case 't': return 3;
    default: return -1;
    }
}

/* Contig.reverse: a<->t, c<->g; anything else has no defined complement -> 'n' (LITERAL policy) */
static inline char complement(char c) {
    switch (c) {
    case 'a': return 't';
    case 'c': return 'g';
    case 'g': return 'c';
    case 't': return 'a';
    default: return 'n';
    }
}

/* ------------------------------------------------------------------------------------------ */
/* STRING mode: the HashSet<String> analogue                                                  */
/* ------------------------------------------------------------------------------------------ */

struct orc_strset {
    int alphabet, k, policy;
    char *text;       /* arena of folded contigs, each followed by its reverse complement */
    size_t text_len, text_cap;
    uint32_t *tab;    /* open addressing; entry = offset+1 into text, 0 = empty */
    size_t cap, n;
};

/* java.lang.String.hashCode over K chars, then java.util.HashMap.hash spread */
static inline uint32_t java_hash(const char *s, int k) {
    uint32_t h = 0;
    for (int i = 0; i < k; i++) h = 31u * h + (uint8_t)s[i];
    return h ^ (h >> 16);
}

orc_strset *orc_strset_new(int alphabet, int k, int ambig_policy) {
    if (k < 1) return NULL;
    orc_strset *s = (orc_strset *)calloc(1, sizeof(*s));
    if (!s) return NULL;
    s->alphabet = alphabet;
    s->k = k;
    s->policy = ambig_policy;
    s->cap = 1024;
    s->tab = (uint32_t *)calloc(s->cap, sizeof(uint32_t));
    if (!s->tab) {
        free(s);
        return NULL;
    }
    return s;
}

void orc_strset_free(orc_strset *s) {
    if (!s) return;
    free(s->text);
    free(s->tab);
    free(s);
}

size_t orc_strset_size(const orc_strset *s) { return s->n; }

static int strset_contains(const orc_strset *s, const char *kmer) {
    size_t mask = s->cap - 1;
    size_t pos = java_hash(kmer, s->k) & mask;
    for (;;) {
        uint32_t e = s->tab[pos];
        if (e == 0) return 0;
        if (memcmp(s->text + (e - 1), kmer, (size_t)s->k) == 0) return 1;
        pos = (pos + 1) & mask;
    }
}

static void strset_place(orc_strset *s, uint32_t off1) {
    size_t mask = s->cap - 1;
    size_t pos = java_hash(s->text + (off1 - 1), s->k) & mask;
    while (s->tab[pos] != 0) pos = (pos + 1) & mask;
    s->tab[pos] = off1;
}

static int strset_reserve(orc_strset *s, size_t want_n) {
    if (want_n * 2 <= s->cap) return 0;
    size_t ncap = s->cap;
    while (want_n * 2 > ncap) ncap <<= 1;
    uint32_t *old = s->tab;
    size_t ocap = s->cap;
    uint32_t *nt = (uint32_t *)calloc(ncap, sizeof(uint32_t));
    if (!nt) return -1;
    s->tab = nt;
    s->cap = ncap;
    for (size_t i = 0; i < ocap; i++)
        if (old[i]) strset_place(s, old[i]);
    free(old);
    return 0;
}

/* HashSet.add: insert the K chars at text offset off unless an equal string is present */
static void strset_insert(orc_strset *s, size_t off) {
    size_t mask = s->cap - 1;
    const char *kmer = s->text + off;
    size_t pos = java_hash(kmer, s->k) & mask;
    for (;;) {
        uint32_t e = s->tab[pos];
        if (e == 0) {
            s->tab[pos] = (uint32_t)(off + 1);
            s->n++;
            return;
        }
        if (memcmp(s->text + (e - 1), kmer, (size_t)s->k) == 0) return;
        pos = (pos + 1) & mask;
    }
}

/* insert every length-K window of text[base, base+len) honouring the ambiguity policy */
static void strset_insert_windows(orc_strset *s, size_t base, size_t len, int check_acgt) {
    size_t k = (size_t)s->k;
    if (len < k) return;
    size_t run = 0; /* length of the current run of acgt characters ending at position p */
    for (size_t p = 0; p < len; p++) {
        if (check_acgt) run = (nuc_code(s->text[base + p]) >= 0) ? run + 1 : 0;
        else run = p + 1;
        if (p + 1 >= k && run >= k) strset_insert(s, base + p + 1 - k);
    }
}

int orc_strset_add(orc_strset *s, const char *seq, size_t len) {
    int nuc = (s->alphabet != ORC_PROT);
    size_t need = s->text_len + (nuc ? 2 * len : len);
    if (need >= 0xFFFFFFF0u) return -1;
    if (need > s->text_cap) {
        size_t nc = s->text_cap ? s->text_cap : 4096;
        while (nc < need) nc <<= 1;
        char *nt = (char *)realloc(s->text, nc);
        if (!nt) return -1;
        s->text = nt;
        s->text_cap = nc;
    }
    size_t k = (size_t)s->k;
    size_t windows = len >= k ? len - k + 1 : 0;
    if (strset_reserve(s, s->n + (nuc ? 2 : 1) * windows)) return -1;
    size_t base = s->text_len;
    if (nuc) {
        /* DnaKmers / GenomeKmers: k-mers of the lower-cased sequence and of its reverse complement */
        for (size_t i = 0; i < len; i++) s->text[base + i] = fold_nuc(seq[i], s->alphabet);
        for (size_t i = 0; i < len; i++) s->text[base + len + i] = complement(s->text[base + len - 1 - i]);
        s->text_len = base + 2 * len;
        int check = (s->policy == ORC_AMBIG_SKIP);
        strset_insert_windows(s, base, len, check);
        strset_insert_windows(s, base + len, len, check);
    } else {
        /* ProteinKmers: literal substrings, single strand, no folding */
        memcpy(s->text + base, seq, len);
        s->text_len = base + len;
        strset_insert_windows(s, base, len, 0);
    }
    return 0;
}

/* SequenceKmers.similarity: count members of b that are present in a */
size_t orc_strset_similarity(const orc_strset *a, const orc_strset *b) {
    if (a->k != b->k) return 0;
    size_t hits = 0;
    for (size_t i = 0; i < b->cap; i++) {
        uint32_t e = b->tab[i];
        if (e && strset_contains(a, b->text + (e - 1))) hits++;
    }
    return hits;
}

/* ------------------------------------------------------------------------------------------ */
/* INTEGER mode: sorted unique canonical keys                                                 */
/* ------------------------------------------------------------------------------------------ */

struct orc_intset {
    int alphabet, k;
    uint64_t *keys;
    size_t n, cap;
    size_t n_pal;
    int finished;
};

uint64_t orc_dna_revcomp_key(uint64_t key, int k) {
    uint64_t rc = 0;
    for (int i = 0; i < k; i++) {
        rc = (rc << 2) | (3u - (key & 3u));
        key >>= 2;
    }
    return rc;
}

uint64_t orc_dna_canonical(const char *kmer, int k, int *valid) {
    uint64_t fwd = 0;
    *valid = 1;
    for (int i = 0; i < k; i++) {
        int c = nuc_code(lower_ascii(kmer[i]));
        if (c < 0) {
            *valid = 0;
            return UINT64_MAX;
        }
        fwd = (fwd << 2) | (uint64_t)c;
    }
    uint64_t rc = orc_dna_revcomp_key(fwd, k);
    return fwd < rc ? fwd : rc;
}

orc_intset *orc_intset_new(int alphabet, int k) {
    if (k < 1) return NULL;
    if (alphabet == ORC_PROT ? k > 8 : k > 32) return NULL;
    orc_intset *s = (orc_intset *)calloc(1, sizeof(*s));
    if (!s) return NULL;
    s->alphabet = alphabet;
    s->k = k;
    return s;
}

void orc_intset_free(orc_intset *s) {
    if (!s) return;
    free(s->keys);
    free(s);
}

static int intset_push(orc_intset *s, uint64_t key) {
    if (s->n == s->cap) {
        size_t nc = s->cap ? s->cap * 2 : 4096;
        uint64_t *nk = (uint64_t *)realloc(s->keys, nc * sizeof(uint64_t));
        if (!nk) return -1;
        s->keys = nk;
        s->cap = nc;
    }
    s->keys[s->n++] = key;
    return 0;
}

int orc_intset_add(orc_intset *s, const char *seq, size_t len) {
    int k = s->k;
    if (s->finished) return -1;
    if (s->alphabet == ORC_PROT) {
        /* K raw bytes, first character most significant: key order == String order */
        uint64_t mask = (k == 8) ? UINT64_MAX : ((1ull << (8 * k)) - 1);
        uint64_t key = 0;
        for (size_t p = 0; p < len; p++) {
            key = ((key << 8) | (uint8_t)seq[p]) & mask;
            if (p + 1 >= (size_t)k && intset_push(s, key)) return -1;
        }
        return 0;
    }
    uint64_t mask = (k == 32) ? UINT64_MAX : ((1ull << (2 * k)) - 1);
    int shift = 2 * (k - 1);
    uint64_t fwd = 0, rc = 0;
    size_t run = 0;
    for (size_t p = 0; p < len; p++) {
        int c = nuc_code(fold_nuc(seq[p], s->alphabet));
        if (c < 0) {
            run = 0;
            fwd = rc = 0;
            continue;
        }
        fwd = ((fwd << 2) | (uint64_t)c) & mask;
        rc = (rc >> 2) | ((uint64_t)(3 - c) << shift);
        run++;
        if (run >= (size_t)k && intset_push(s, fwd < rc ? fwd : rc)) return -1;
    }
    return 0;
}

static int cmp_u64(const void *x, const void *y) {
    uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
    return a < b ? -1 : (a > b ? 1 : 0);
}

static void radix_sort_u64(uint64_t *a, size_t n) {
    if (n < 2) return;
    uint64_t *tmp = (uint64_t *)malloc(n * sizeof(uint64_t));
    if (!tmp) { /* fall back to an in-place O(n log n) sort */
        qsort(a, n, sizeof(uint64_t), cmp_u64);
        return;
    }
    uint64_t *src = a, *dst = tmp;
    for (int pass = 0; pass < 8; pass++) {
        size_t hist[256];
        memset(hist, 0, sizeof(hist));
        int sh = pass * 8;
        for (size_t i = 0; i < n; i++) hist[(src[i] >> sh) & 0xFF]++;
        int trivial = 0;
        for (int d = 0; d < 256; d++)
            if (hist[d] == n) trivial = 1;
        if (trivial) continue;
        size_t sum = 0;
        for (int d = 0; d < 256; d++) {
            size_t c = hist[d];
            hist[d] = sum;
            sum += c;
        }
        for (size_t i = 0; i < n; i++) dst[hist[(src[i] >> sh) & 0xFF]++] = src[i];
        uint64_t *t = src;
        src = dst;
        dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(uint64_t));
    free(tmp);
}

int orc_intset_finish(orc_intset *s) {
    if (s->finished) return 0;
    radix_sort_u64(s->keys, s->n);
    size_t w = 0;
    for (size_t i = 0; i < s->n; i++)
        if (w == 0 || s->keys[i] != s->keys[w - 1]) s->keys[w++] = s->keys[i];
    s->n = w;
    s->n_pal = 0;
    if (s->alphabet != ORC_PROT && (s->k % 2) == 0)
        for (size_t i = 0; i < s->n; i++)
            if (orc_dna_revcomp_key(s->keys[i], s->k) == s->keys[i]) s->n_pal++;
    s->finished = 1;
    return 0;
}

size_t orc_intset_count(const orc_intset *s) { return s->n; }
size_t orc_intset_palindromes(const orc_intset *s) { return s->n_pal; }
size_t orc_intset_size_both(const orc_intset *s) {
    return s->alphabet == ORC_PROT ? s->n : 2 * s->n - s->n_pal;
}
const uint64_t *orc_intset_keys(const orc_intset *s) { return s->keys; }

size_t orc_intset_intersect(const orc_intset *a, const orc_intset *b, size_t *pal_inter) {
    size_t i = 0, j = 0, hits = 0, pal = 0;
    int check_pal = (a->alphabet != ORC_PROT) && (a->k % 2 == 0);
    while (i < a->n && j < b->n) {
        uint64_t x = a->keys[i], y = b->keys[j];
        if (x < y) i++;
        else if (y < x) j++;
        else {
            hits++;
            if (check_pal && orc_dna_revcomp_key(x, a->k) == x) pal++;
            i++;
            j++;
        }
    }
    if (pal_inter) *pal_inter = pal;
    return hits;
}

/* ------------------------------------------------------------------------------------------ */
/* formula + Java text                                                                        */
/* ------------------------------------------------------------------------------------------ */

double orc_distance(uint64_t inter, uint64_t size_a, uint64_t size_b) {
    double ret = 1.0;
    double similarity = (double)inter;
    if (similarity > 0) {
        /* (this.size() + other.size()) is a Java int addition (wraps at 2^31) */
        int32_t sum = (int32_t)((uint32_t)size_a + (uint32_t)size_b);
        double uni = (double)sum - similarity;
        ret = 1.0 - similarity / uni;
    }
    return ret;
}

int orc_double_to_string(double v, char *buf, size_t cap) {
    if (isnan(v)) return snprintf(buf, cap, "NaN");
    if (isinf(v)) return snprintf(buf, cap, v > 0 ? "Infinity" : "-Infinity");
    if (v == 0.0) return snprintf(buf, cap, signbit(v) ? "-0.0" : "0.0");
    /* shortest decimal that round-trips (what JDK >= 19 prints) */
    char sci[40];
    int prec;
    for (prec = 1; prec <= 17; prec++) {
        snprintf(sci, sizeof(sci), "%.*e", prec - 1, v);
        if (strtod(sci, NULL) == v) break;
    }
    /* sci = [-]d[.ddd]e[+-]xx */
    char digits[24];
    int nd = 0, neg = 0;
    const char *p = sci;
    if (*p == '-') {
        neg = 1;
        p++;
    }
    for (; *p && *p != 'e'; p++)
        if (*p >= '0' && *p <= '9') digits[nd++] = *p;
    int exp10 = atoi(p + 1);
    while (nd > 1 && digits[nd - 1] == '0') nd--;
    digits[nd] = 0;
    char out[64];
    int o = 0;
    if (neg) out[o++] = '-';
    if (exp10 >= -3 && exp10 < 7) {
        if (exp10 >= 0) {
            for (int i = 0; i <= exp10; i++) out[o++] = i < nd ? digits[i] : '0';
            out[o++] = '.';
            if (nd > exp10 + 1)
                for (int i = exp10 + 1; i < nd; i++) out[o++] = digits[i];
            else out[o++] = '0';
        } else {
            out[o++] = '0';
            out[o++] = '.';
            for (int i = 0; i < -exp10 - 1; i++) out[o++] = '0';
            for (int i = 0; i < nd; i++) out[o++] = digits[i];
        }
    } else {
        out[o++] = digits[0];
        out[o++] = '.';
        if (nd > 1)
            for (int i = 1; i < nd; i++) out[o++] = digits[i];
        else out[o++] = '0';
        o += snprintf(out + o, sizeof(out) - (size_t)o, "E%d", exp10);
    }
    out[o] = 0;
    return snprintf(buf, cap, "%s", out);
}

/* ------------------------------------------------------------------------------------------ */
/* whole-command restatements (timed CPU baseline)                                            */
/* ------------------------------------------------------------------------------------------ */

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

/* minimal dynamic parallel-for (the reference uses IntStream.parallel() on the common pool) */
typedef struct {
    void (*body)(long idx, void *arg);
    void *arg;
    long n;
    long next;
} pfor_t;

static void *pfor_worker(void *p) {
    pfor_t *pf = (pfor_t *)p;
    for (;;) {
        long i = __atomic_fetch_add(&pf->next, 1, __ATOMIC_RELAXED);
        if (i >= pf->n) break;
        pf->body(i, pf->arg);
    }
    return NULL;
}

static void parallel_for(long n, int threads, void (*body)(long, void *), void *arg) {
    pfor_t pf = {body, arg, n, 0};
    if (threads > n) threads = (int)n;
    if (threads <= 1) {
        pfor_worker(&pf);
        return;
    }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    int started = 0;
    if (tid)
        for (int t = 1; t < threads; t++)
            if (pthread_create(&tid[started], NULL, pfor_worker, &pf) == 0) started++;
    pfor_worker(&pf);
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
    free(tid);
}

typedef struct {
    int mode;
    orc_strset *s;
    orc_intset *i;
} anyset;

/* KmerType.createKmers(seq, K) */
static int anyset_build(anyset *out, int mode, int alphabet, int k, int ambig, const char *seq, size_t len) {
    out->mode = mode;
    out->s = NULL;
    out->i = NULL;
    if (mode == 0) {
        out->s = orc_strset_new(alphabet, k, ambig);
        if (!out->s || orc_strset_add(out->s, seq, len)) return -1;
    } else {
        out->i = orc_intset_new(alphabet, k);
        if (!out->i || orc_intset_add(out->i, seq, len) || orc_intset_finish(out->i)) return -1;
    }
    return 0;
}

static void anyset_free(anyset *a) {
    orc_strset_free(a->s);
    orc_intset_free(a->i);
    a->s = NULL;
    a->i = NULL;
}

/* SequenceKmers.distance: similarity + formula on the both-strand sizes */
static void anyset_distance(const anyset *a, const anyset *b, uint64_t *inter, double *dist) {
    uint64_t I, sa, sb;
    if (a->mode == 0) {
        I = orc_strset_similarity(a->s, b->s);
        sa = orc_strset_size(a->s);
        sb = orc_strset_size(b->s);
    } else {
        size_t pal = 0;
        size_t c = orc_intset_intersect(a->i, b->i, &pal);
        I = (a->i->alphabet == ORC_PROT) ? c : 2 * c - pal;
        sa = orc_intset_size_both(a->i);
        sb = orc_intset_size_both(b->i);
    }
    *inter = I;
    *dist = orc_distance(I, sa, sb);
}

static inline size_t tri_index(size_t n, size_t i, size_t j) { /* i < j */
    return i * (2 * n - i - 1) / 2 + (j - i - 1);
}

typedef struct {
    const char *const *seqs;
    const size_t *lens;
    size_t n, b0;
    int alphabet, k, batch, mode, ambig;
    anyset *cache;
    uint64_t *inter;
    double *dist;
    int rc;
} fd_job;

/* FastaDistanceProcessor.computePairs (:174-194) for row b0+ii of the current batch */
static void fd_row(long ii, void *arg) {
    fd_job *job = (fd_job *)arg;
    size_t i = job->b0 + (size_t)ii;
    for (size_t j = i + 1; j < job->n; j++) { /* :177 */
        anyset tmp;
        const anyset *other;
        int built = 0;
        if (j - job->b0 < (size_t)job->batch) other = &job->cache[j - job->b0]; /* :181-182 */
        else { /* :183-184 -- the set of an uncached column is rebuilt for every (row, column) */
            if (anyset_build(&tmp, job->mode, job->alphabet, job->k, job->ambig, job->seqs[j], job->lens[j])) {
                anyset_free(&tmp);
                __atomic_store_n(&job->rc, -1, __ATOMIC_RELAXED);
                continue;
            }
            other = &tmp;
            built = 1;
        }
        size_t t = tri_index(job->n, i, j);
        anyset_distance(&job->cache[ii], other, &job->inter[t], &job->dist[t]); /* :186 */
        if (built) anyset_free(&tmp);
    }
}

int orc_fasta_dist(const char *const *seqs, const size_t *lens, size_t n, int alphabet, int k,
                   int batch, int threads, int mode, int ambig, uint64_t *inter, double *dist) {
    if (batch < 1) return -1;
    if (mode != 0 && ambig != ORC_AMBIG_SKIP) return -2; /* integer keys cannot hold literal k-mers */
    if (threads < 1) threads = orc_max_threads();
    anyset *cache = (anyset *)calloc((size_t)batch, sizeof(anyset));
    if (!cache) return -1;
    fd_job job = {seqs, lens, n, 0, alphabet, k, batch, mode, ambig, cache, inter, dist, 0};
    /* FastaDistanceProcessor.java:141 -- loop until the sequence list is empty */
    for (size_t b0 = 0; b0 < n && job.rc == 0; b0 += (size_t)batch) {
        size_t bsize = (n - b0 < (size_t)batch) ? n - b0 : (size_t)batch; /* :145-149 */
        for (size_t i = 0; i < bsize; i++)                                /* :151-155, serial */
            if (anyset_build(&cache[i], mode, alphabet, k, ambig, seqs[b0 + i], lens[b0 + i])) job.rc = -1;
        job.b0 = b0;
        if (job.rc == 0) parallel_for((long)bsize, threads, fd_row, &job); /* :157-158 */
        for (size_t i = 0; i < bsize; i++) anyset_free(&cache[i]);
    }
    free(cache);
    return job.rc;
}

typedef struct {
    const anyset *qs;
    const anyset *refs;
    uint64_t *inter;
    double *dist;
} qr_job;

static void qr_one(long i, void *arg) {
    qr_job *job = (qr_job *)arg;
    anyset_distance(job->qs, &job->refs[i], &job->inter[i], &job->dist[i]); /* GenomeProcessor.java:140 */
}

int orc_query_vs_ref(const char *const *q, const size_t *qlens, size_t nq, const char *const *r,
                     const size_t *rlens, size_t nr, int alphabet, int k, int threads, int mode, int ambig,
                     uint64_t *inter, double *dist) {
    if (mode != 0 && ambig != ORC_AMBIG_SKIP) return -2;
    if (threads < 1) threads = orc_max_threads();
    anyset *refs = (anyset *)calloc(nr ? nr : 1, sizeof(anyset));
    if (!refs) return -1;
    int rc = 0;
    /* GenomeProcessor.java:101-111 -- base genomes loaded serially, all resident */
    for (size_t i = 0; i < nr; i++)
        if (anyset_build(&refs[i], mode, alphabet, k, ambig, r[i], rlens[i])) rc = -1;
    /* :129-147 -- one query at a time against all base genomes in parallel */
    for (size_t qi = 0; qi < nq && rc == 0; qi++) {
        anyset qs;
        if (anyset_build(&qs, mode, alphabet, k, ambig, q[qi], qlens[qi])) {
            anyset_free(&qs);
            rc = -1;
            break;
        }
        qr_job job = {&qs, refs, inter + qi * nr, dist + qi * nr};
        parallel_for((long)nr, threads, qr_one, &job);
        anyset_free(&qs);
    }
    for (size_t i = 0; i < nr; i++) anyset_free(&refs[i]);
    free(refs);
    return rc;
}
