package org.theseed.sequence.gpu;

import java.io.File;
import java.io.IOException;
import java.nio.charset.StandardCharsets;
import java.util.List;

import org.theseed.basic.ParseFailureException;

/**
 * Batched GPU replacement for the k-mer objects of org.theseed.sequence: the processors hand it whole
 * sequence lists and receive distance blocks, instead of calling SequenceKmers.distance() per pair.
 *
 * K is per engine (the reference keeps it in the statics GenomeKmers.setKmerSize /
 * ProteinKmers.setKmerSize, GenomeProcessor.java:86, ProteinKmerReader.java:92).
 */
public class GpuKmerEngine implements AutoCloseable {

    /** KmerType ordinal mapping used by the native side */
    public static final int DNA = 0, PROT = 1, RNA = 2;

    private long ctx;

    public GpuKmerEngine(int alphabet, int kmerSize, int device) throws ParseFailureException {
        this.ctx = GkdNative.create(device, kmerSize, alphabet, 0);
        if (this.ctx == 0)
            throw new ParseFailureException(GkdNative.lastError(0));
    }

    /** KmerType.createKmers(seq, K) / new ProteinKmers(str): returns the handle of the new set */
    public int add(String sequence) {
        return check(GkdNative.addSequences(this.ctx, new byte[][] { sequence.getBytes(StandardCharsets.ISO_8859_1) }));
    }

    /** new GenomeKmers(genome): one piece per contig; k-mers never span contigs */
    public int addGenome(List<String> contigs) {
        byte[][] parts = new byte[contigs.size()][];
        for (int i = 0; i < parts.length; i++)
            parts[i] = contigs.get(i).getBytes(StandardCharsets.ISO_8859_1);
        return check(GkdNative.addSequences(this.ctx, parts));
    }

    /** FastaInputStream(File): every record becomes a set; returns {firstId, count} */
    public int[] addFasta(File inFile) throws IOException {
        int[] r = GkdNative.addFastaFile(this.ctx, inFile.getPath(), true);
        if (r == null)
            throw new IOException(GkdNative.lastError(this.ctx));
        return r;
    }

    public String getLabel(int id) { return GkdNative.label(this.ctx, id); }
    public String getComment(int id) { return GkdNative.comment(this.ctx, id); }

    public void build() { check(GkdNative.buildSets(this.ctx)); }

    /** all pairs i &lt; j in list order, row-major strict upper triangle */
    public double[] allVsAll(int n) {
        double[] dist = new double[n * (n - 1) / 2];
        check(GkdNative.allVsAll(this.ctx, dist));
        return dist;
    }

    /** every query against every reference, row-major */
    public double[] queryVsRef(int[] queries, int[] refs) {
        double[] dist = new double[queries.length * refs.length];
        check(GkdNative.queryVsRef(this.ctx, queries, refs, dist));
        return dist;
    }

    /** SequenceKmers.distance(other) for the greedy callers (DistanceRepsProcessor, FastaDistanceRepsProcessor) */
    public double distance(int a, int b) { return GkdNative.pairDistance(this.ctx, a, b); }

    private int check(int rc) {
        if (rc < 0) {
            String msg = GkdNative.lastError(this.ctx);
            if (rc == -1) throw new IllegalArgumentException(msg);   // GKD_EINVAL
            throw new RuntimeException(msg);                          // GKD_ENOMEM / GKD_ECUDA / GKD_ESTATE
        }
        return rc;
    }

    @Override
    public void close() {
        if (this.ctx != 0) {
            GkdNative.destroy(this.ctx);
            this.ctx = 0;
        }
    }
}
