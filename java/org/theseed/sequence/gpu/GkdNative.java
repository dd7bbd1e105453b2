package org.theseed.sequence.gpu;

/**
 * JNI declarations for libgkd.so (include/gkd.h).  One native method per C entry point the
 * replacement processors need; handles are opaque longs, errors come back as negative status codes
 * and are turned into exceptions by {@link GpuKmerEngine}.
 *
 * Not compiled in the build image (no JDK there); see INTEGRATION.md for the build line.
 */
final class GkdNative {
    static {
        System.loadLibrary("gkd_jni"); // links against libgkd.so
    }

    private GkdNative() { }

    static native long create(int device, int k, int alphabet, int strandMode);
    static native void destroy(long ctx);
    static native String lastError(long ctx);
    /** one genome / record from contigs (Latin-1 bytes); returns the set id or a negative status */
    static native int addSequences(long ctx, byte[][] contigs);
    /** every record of a FASTA file; returns {firstId, count} or null on error */
    static native int[] addFastaFile(long ctx, String path, boolean perRecord);
    static native String label(long ctx, int id);
    static native String comment(long ctx, int id);
    static native int buildSets(long ctx);
    /** fills dist (length n*(n-1)/2, row-major strict upper triangle) */
    static native int allVsAll(long ctx, double[] dist);
    /** fills dist (length q.length * r.length, row-major) */
    static native int queryVsRef(long ctx, int[] q, int[] r, double[] dist);
    /** SequenceKmers.distance for one pair */
    static native double pairDistance(long ctx, int a, int b);
    static native long setSize(long ctx, int id);
}
