/*
 * gkd_oracle.h -- CPU ORACLE for the k-mer set distance hot path of SEEDtk/genome.distance.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path (libgkd.so) never
 * links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED: the reference tree (/root/reference) ships no tests, fixtures or golden vectors,
 * and the classes that hold the arithmetic (org.theseed.sequence.{SequenceKmers,DnaKmers,GenomeKmers,
 * ProteinKmers,KmerType}, Maven artifact org.theseed:sequence:1.0.0, pom.xml:45-49) are not vendored
 * and no JVM exists in the build image.  This file restates the published algorithm of that artifact
 * as it is observable from the reference's call sites:
 *   - k-mers are literal String substrings held in a HashSet<String>
 *       (KmerCountProcessor.java:76-77 iterates a ProteinKmers as Strings)
 *   - DNA sets hold the k-mers of the lower-cased sequence AND of its reverse complement
 *       (KmerType.DNA.createKmers, FastaDistanceProcessor.java:153,184; GenomeKmers per contig,
 *        GenomeProcessor.java:109,139)
 *   - protein sets hold the k-mers of the single strand (ProteinKmerReader.java:100-101)
 *   - similarity = number of members of one set found by probing the other
 *   - distance = I==0 ? 1.0 : 1.0 - I / ((|A|+|B|) - I)   in double, |A|+|B| added as Java int
 *       (FastaDistanceProcessor.java:186, GenomeProcessor.java:140; 1.0 special value confirmed by
 *        DistanceRepsProcessor.java:108-111, GroupTypeSpec.java:84,90)
 * The known-answer vectors it is checked against are the hand-derived ones of SURVEY.md section 8(c).
 *
 * Two independent restatements live here and are cross-checked by the tests:
 *   STRING mode  - hash set of literal K-character substrings, Java String.hashCode + HashMap spread;
 *                  this is the faithful analogue of the reference's HashSet<String> and the timed
 *                  CPU baseline ("port").
 *   INTEGER mode - sorted unique canonical 2-bit (DNA) / raw-byte (protein) uint64 keys plus the
 *                  palindrome count, i.e. the representation the CUDA path uses; related to STRING
 *                  mode by |S| = 2|C| - P and |S_A n S_B| = 2|C_A n C_B| - P(C_A n C_B).
 */
#ifndef GKD_ORACLE_H
#define GKD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sequence alphabets (KmerType members; RNA folds u->t and is otherwise DNA) */
enum { ORC_DNA = 0, ORC_PROT = 1, ORC_RNA = 2 };
/* what to do with a DNA k-mer that contains a character outside acgt (unpinned upstream):
 *   ORC_AMBIG_SKIP    - the k-mer (and its reverse complement) is not inserted        [engine default]
 *   ORC_AMBIG_LITERAL - kept as a literal string; the complement of an unknown base is 'n'
 *                       (STRING mode only; INTEGER mode cannot represent it)                      */
enum { ORC_AMBIG_SKIP = 0, ORC_AMBIG_LITERAL = 1 };

/* ---------------- STRING mode (HashSet<String> analogue) ---------------- */
typedef struct orc_strset orc_strset;
orc_strset *orc_strset_new(int alphabet, int k, int ambig_policy);
/* add every k-mer of one contig / FASTA record / protein; k-mers never span calls */
int orc_strset_add(orc_strset *s, const char *seq, size_t len);
size_t orc_strset_size(const orc_strset *s);
/* |A n B| by probing A with every member of B (SequenceKmers.similarity) */
size_t orc_strset_similarity(const orc_strset *a, const orc_strset *b);
void orc_strset_free(orc_strset *s);

/* ---------------- INTEGER mode (sorted canonical uint64 keys) ---------------- */
typedef struct orc_intset orc_intset;
orc_intset *orc_intset_new(int alphabet, int k);
int orc_intset_add(orc_intset *s, const char *seq, size_t len);
/* sort + unique; must be called once after the last add */
int orc_intset_finish(orc_intset *s);
size_t orc_intset_count(const orc_intset *s);        /* distinct canonical keys |C| */
size_t orc_intset_palindromes(const orc_intset *s);  /* members equal to their own reverse complement */
size_t orc_intset_size_both(const orc_intset *s);    /* 2|C| - P for DNA, |C| for protein */
const uint64_t *orc_intset_keys(const orc_intset *s);
/* linear merge; *pal_inter receives the number of palindromic members of the intersection */
size_t orc_intset_intersect(const orc_intset *a, const orc_intset *b, size_t *pal_inter);
void orc_intset_free(orc_intset *s);

/* key helpers (exposed so the tests can pin the encoding) */
uint64_t orc_dna_canonical(const char *kmer, int k, int *valid); /* min(fwd, revcomp), a=0 c=1 g=2 t=3 */
uint64_t orc_dna_revcomp_key(uint64_t key, int k);

/* ---------------- formula + text ---------------- */
/* Java: double sim = I; if (sim > 0) { double u = (sizeA + sizeB) - sim; r = 1.0 - sim / u; } */
double orc_distance(uint64_t inter, uint64_t size_a, uint64_t size_b);
/* java.lang.Double.toString (JDK >= 19 shortest-repr); returns length written (excluding NUL) */
int orc_double_to_string(double v, char *buf, size_t cap);

/* ---------------- whole-command restatements (used as the timed CPU baseline) ---------------- */
/* FastaDistanceProcessor.runReporter (:134-165) + computePairs (:174-194): all pairs i<j of n
 * records, batch-cached sets, uncached columns rebuilt per (row, column), rows in parallel.
 * inter/dist are row-major strict upper triangle, length n*(n-1)/2.  mode: 0 = STRING, 1 = INTEGER.
 * ambig = ORC_AMBIG_* (the literal policy exists in STRING mode only).
 * Returns 0, -1 on allocation failure, -2 for INTEGER mode with the literal policy. */
int orc_fasta_dist(const char *const *seqs, const size_t *lens, size_t n, int alphabet, int k,
                   int batch, int threads, int mode, int ambig, uint64_t *inter, double *dist);
/* GenomeProcessor.runReporter (:119-150): every query against every base genome; one record per
 * genome here (multi-contig genomes go through the set API).  inter/dist are nq*nr row-major. */
int orc_query_vs_ref(const char *const *q, const size_t *qlens, size_t nq, const char *const *r,
                     const size_t *rlens, size_t nr, int alphabet, int k, int threads, int mode, int ambig,
                     uint64_t *inter, double *dist);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
