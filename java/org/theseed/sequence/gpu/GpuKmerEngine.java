package org.theseed.sequence.gpu;

import java.io.File;
import java.io.IOException;
import java.nio.charset.StandardCharsets;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

import org.theseed.basic.ParseFailureException;

/**
 * Batched GPU replacement for the k-mer objects of org.theseed.sequence: the processors hand it whole
 * sequence lists and receive distance blocks, instead of calling SequenceKmers.distance() per pair.
 * The drop-in classes org.theseed.sequence.{KmerType,SequenceKmers,GenomeKmers,ProteinKmers} are handles
 * into the engine returned by {@link #shared(int, int)}.
 *
 * K is per engine (the reference keeps it in the statics GenomeKmers.setKmerSize /
 * ProteinKmers.setKmerSize, GenomeProcessor.java:86, ProteinKmerReader.java:92).
 */
public class GpuKmerEngine implements AutoCloseable {

    /** KmerType ordinal mapping used by the native side */
    public static final int DNA = 0, PROT = 1, RNA = 2;

    /** one engine per (alphabet, K) for the object-style API, so objects of one kind can be compared */
    private static final Map<Long, GpuKmerEngine> SHARED = new HashMap<Long, GpuKmerEngine>();

    private long ctx;
    /** sets added since the last build (built lazily by the first distance request) */
    private boolean dirty;

    public GpuKmerEngine(int alphabet, int kmerSize, int device) throws ParseFailureException {
        this.ctx = GkdNative.create(device, kmerSize, alphabet, 0);
        if (this.ctx == 0)
            throw new ParseFailureException(GkdNative.lastError(0));
        this.dirty = false;
    }

    /** @return the process-wide engine for this alphabet and kmer size (device 0) */
    public static synchronized GpuKmerEngine shared(int alphabet, int kmerSize) {
        long key = ((long) alphabet << 32) | (kmerSize & 0xFFFFFFFFL);
        GpuKmerEngine retVal = SHARED.get(key);
        if (retVal == null) {
            try {
                retVal = new GpuKmerEngine(alphabet, kmerSize, 0);
            } catch (ParseFailureException e) {
                throw new IllegalArgumentException(e.getMessage());
            }
            SHARED.put(key, retVal);
        }
        return retVal;
    }

    /** KmerType.createKmers(seq, K) / new ProteinKmers(str): returns the handle of the new set */
    public synchronized int add(String sequence) {
        this.dirty = true;
        return check(GkdNative.addSequences(this.ctx, new byte[][] { sequence.getBytes(StandardCharsets.ISO_8859_1) }));
    }

    /** new GenomeKmers(genome): one piece per contig; k-mers never span contigs */
    public synchronized int addGenome(List<String> contigs) {
        byte[][] parts = new byte[contigs.size()][];
        for (int i = 0; i < parts.length; i++)
            parts[i] = contigs.get(i).getBytes(StandardCharsets.ISO_8859_1);
        this.dirty = true;
        return check(GkdNative.addSequences(this.ctx, parts));
    }

    /** FastaInputStream(File): every record becomes a set; returns {firstId, count} */
    public synchronized int[] addFasta(File inFile) throws IOException {
        int[] r = GkdNative.addFastaFile(this.ctx, inFile.getPath(), true);
        if (r == null)
            throw new IOException(GkdNative.lastError(this.ctx));
        this.dirty = true;
        return r;
    }

    public String getLabel(int id) { return GkdNative.label(this.ctx, id); }
    public String getComment(int id) { return GkdNative.comment(this.ctx, id); }

    /** @return the number of sets in the engine */
    public int size() { return GkdNative.count(this.ctx); }

    /** build the sets of everything added since the last build (kernels 1-3) */
    public synchronized void build() {
        if (this.dirty) {
            check(GkdNative.buildSets(this.ctx));
            this.dirty = false;
        }
    }

    /** drop the sets with id &gt;= keep (e.g. query genomes that have been reported) */
    public synchronized void truncate(int keep) { check(GkdNative.truncate(this.ctx, keep)); }

    /** all pairs i &lt; j in list order, row-major strict upper triangle */
    public synchronized double[] allVsAll() {
        this.build();
        long n = this.size();
        long pairs = n * (n - 1) / 2;          // long arithmetic: n can exceed 65,536
        if (pairs > Integer.MAX_VALUE - 8)
            throw new IllegalArgumentException("Too many pairs for one array; use allVsAllRange in blocks.");
        double[] dist = new double[(int) pairs];
        check(GkdNative.allVsAllRange(this.ctx, (int) n, 0L, pairs, dist));
        return dist;
    }

    /** pairs [first, first + count) of the same enumeration, for reports larger than one Java array */
    public synchronized double[] allVsAllRange(long first, int count) {
        this.build();
        double[] dist = new double[count];
        check(GkdNative.allVsAllRange(this.ctx, this.size(), first, (long) count, dist));
        return dist;
    }

    /** every query against every reference, row-major */
    public synchronized double[] queryVsRef(int[] queries, int[] refs) {
        this.build();
        long cells = (long) queries.length * (long) refs.length;
        if (cells > Integer.MAX_VALUE - 8)
            throw new IllegalArgumentException("Too many pairs for one array; split the queries.");
        double[] dist = new double[(int) cells];
        check(GkdNative.queryVsRef(this.ctx, queries, refs, dist));
        return dist;
    }

    /** SequenceKmers.distance(other) for the greedy callers (DistanceRepsProcessor, FastaDistanceRepsProcessor) */
    public synchronized double distance(int a, int b) {
        this.build();
        double[] out = new double[1];
        check(GkdNative.pair(this.ctx, a, b, null, out));
        return out[0];
    }

    /** SequenceKmers.similarity(other): kmers in common */
    public synchronized long similarity(int a, int b) {
        this.build();
        long[] inter = new long[1];
        check(GkdNative.pair(this.ctx, a, b, inter, null));
        return inter[0];
    }

    /** @return the reference's HashSet size of a set (both strands for DNA) */
    public synchronized long setSize(int id) {
        this.build();
        long[] out = new long[1];
        check(GkdNative.setSize(this.ctx, id, out));
        return out[0];
    }

    private int check(int rc) {
        if (rc < 0) {
            String msg = GkdNative.lastError(this.ctx);
            if (rc == -1) throw new IllegalArgumentException(msg);   // GKD_EINVAL
            throw new RuntimeException(msg);                          // GKD_ENOMEM / GKD_ECUDA / GKD_ESTATE
        }
        return rc;
    }

    @Override
    public synchronized void close() {
        if (this.ctx != 0) {
            GkdNative.destroy(this.ctx);
            this.ctx = 0;
        }
    }
}
