package org.theseed.sequence;

import org.theseed.sequence.gpu.GpuKmerEngine;

/**
 * Drop-in for org.theseed.sequence.ProteinKmers (reference call sites: constructor ProteinKmerReader.java:100-101,
 * static setKmerSize ProteinKmerReader.java:92).  Single strand, no alphabet filtering, K &lt;= 8 (the key is the
 * K raw bytes, exact).
 */
public class ProteinKmers extends SequenceKmers {

    private static int kmerSize = 8;

    public static void setKmerSize(int newSize) {
        kmerSize = newSize;
    }

    public static int kmerSize() {
        return kmerSize;
    }

    public ProteinKmers(String protein) {
        super(GpuKmerEngine.shared(GpuKmerEngine.PROT, kmerSize), GpuKmerEngine.shared(GpuKmerEngine.PROT, kmerSize).add(protein));
    }
}
