package org.theseed.sequence;

import org.theseed.sequence.gpu.GpuKmerEngine;

/**
 * Drop-in for org.theseed.sequence.SequenceKmers as the reference uses it: an object you can ask
 * {@code distance(other)} (FastaDistanceProcessor.java:186, GenomeProcessor.java:140,
 * DistanceRepsProcessor.java:101,190, FastaDistanceRepsProcessor.java:128, WidthProcessor.java:161).
 * It is a handle (engine, set id) into device-resident sorted k-mer sets; distance() is
 * I == 0 ? 1.0 : 1.0 - I / ((|A| + |B|) - I) computed by kernel 5 of libgkd.so with the reference's
 * both-strand set sizes.
 *
 * Per-pair calls work (greedy callers) but cost one launch each; bulk callers should use
 * {@link #distances(SequenceKmers, SequenceKmers[])} or the replacement processors, which batch.
 * Not carried over: iteration over the k-mer strings and hashSet(width) (sketch path, out of scope).
 */
public class SequenceKmers {

    protected final GpuKmerEngine engine;
    protected final int id;

    public SequenceKmers(GpuKmerEngine engine, int id) {
        this.engine = engine;
        this.id = id;
    }

    /** @return the kmer distance to another sequence's set (1.0 when nothing is shared) */
    public double distance(SequenceKmers other) {
        this.requireSameEngine(other);
        return this.engine.distance(this.id, other.id);
    }

    /** @return the number of kmers in common (SequenceKmers.similarity) */
    public long similarity(SequenceKmers other) {
        this.requireSameEngine(other);
        return this.engine.similarity(this.id, other.id);
    }

    /** @return the number of kmers in this set (both strands for DNA, like the reference's HashSet) */
    public long size() {
        return this.engine.setSize(this.id);
    }

    /** one query against many: a single batched launch (GenomeProcessor.java:140 does this in a parallel stream) */
    public static double[] distances(SequenceKmers query, SequenceKmers[] others) {
        int[] refs = new int[others.length];
        for (int i = 0; i < others.length; i++) {
            query.requireSameEngine(others[i]);
            refs[i] = others[i].id;
        }
        return query.engine.queryVsRef(new int[] { query.id }, refs);
    }

    public int getId() {
        return this.id;
    }

    public GpuKmerEngine getEngine() {
        return this.engine;
    }

    private void requireSameEngine(SequenceKmers other) {
        if (other.engine != this.engine)
            throw new IllegalArgumentException("Kmer objects of different types or kmer sizes cannot be compared.");
    }
}
