package org.theseed.sequence.gpu;

/**
 * JNI declarations for libgkd.so (include/gkd.h).  One native method per C entry point the
 * replacement processors need; handles are opaque longs, errors come back as negative status codes
 * (gkd_status) and are turned into exceptions by {@link GpuKmerEngine}.  Every method that fills a
 * Java array checks the array length against what the native call will write and returns GKD_EINVAL
 * (-1) on a mismatch.
 *
 * Not compiled in the build image (no JDK there; java/jni/gkd_jni.c is syntax-checked against a stub
 * jni.h by tests/test_abi_cpu.py); see INTEGRATION.md for the build line.
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
    /** number of sets in the context */
    static native int count(long ctx);
    static native int buildSets(long ctx);
    static native int truncate(long ctx, int keep);
    /** fills dist (length count) with pairs [first, first+count) of the row-major strict upper triangle of n sets */
    static native int allVsAllRange(long ctx, int n, long first, long count, double[] dist);
    /** fills dist (length q.length * r.length, row-major) */
    static native int queryVsRef(long ctx, int[] q, int[] r, double[] dist);
    /** SequenceKmers.similarity / distance for one pair; either output may be null (length-1 arrays) */
    static native int pair(long ctx, int a, int b, long[] inter, double[] dist);
    /** out[0] = the reference's HashSet size of set id */
    static native int setSize(long ctx, int id, long[] out);
}
