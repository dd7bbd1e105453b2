package org.theseed.sequence;

import org.theseed.sequence.gpu.GpuKmerEngine;

/**
 * Drop-in for org.theseed.sequence.KmerType (reference call sites: FastaDistanceProcessor.java:43,82,89,
 * 95-96,153,184; FastaDistanceRepsProcessor.java:68,74,82,122).  Same surface -- getKmerSize() and
 * createKmers(String, int) -- but the k-mer objects it creates are handles into a shared GPU engine
 * ({@link GpuKmerEngine}, libgkd.so) instead of HashSet&lt;String&gt; containers.
 *
 * Only the constant DNA is visible in the reference tree; PROT and RNA are named after the upstream
 * javadoc ("21 for DNA or RNA, 8 for proteins") and may need renaming to match the real enum.
 */
public enum KmerType {
    DNA(21, GpuKmerEngine.DNA), PROT(8, GpuKmerEngine.PROT), RNA(21, GpuKmerEngine.RNA);

    private final int kmerSize;
    private final int alphabet;

    KmerType(int kmerSize, int alphabet) {
        this.kmerSize = kmerSize;
        this.alphabet = alphabet;
    }

    /** @return the default kmer size for this sequence type */
    public int getKmerSize() {
        return this.kmerSize;
    }

    /** @return the native alphabet code (gkd_alphabet) */
    public int getAlphabet() {
        return this.alphabet;
    }

    /**
     * @return a k-mer object for one sequence; the set is built lazily, with every other object created
     * since the last build, the first time a distance is asked for (one batched launch instead of one
     * HashSet fill per object)
     */
    public SequenceKmers createKmers(String sequence, int kSize) {
        GpuKmerEngine engine = GpuKmerEngine.shared(this.alphabet, kSize);
        return new SequenceKmers(engine, engine.add(sequence));
    }
}
