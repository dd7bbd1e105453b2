package org.theseed.sequence;

import java.util.ArrayList;
import java.util.List;

import org.theseed.genome.Contig;
import org.theseed.genome.Genome;
import org.theseed.sequence.gpu.GpuKmerEngine;

/**
 * Drop-in for org.theseed.sequence.GenomeKmers (reference call sites: constructor GenomeProcessor.java:109,139,
 * DistanceRepsProcessor.java:188,236; static setKmerSize GenomeProcessor.java:86, DistanceRepsProcessor.java:158;
 * getGenomeId/getGenomeName GenomeProcessor.java:144, DistanceRepsProcessor.java:231,244,249-250).
 * One piece per contig, so k-mers never span contigs; both strands are represented by canonical keys on
 * the device and the reference's both-strand set sizes are recovered exactly (DESIGN.md section 2).
 */
public class GenomeKmers extends SequenceKmers {

    /** kmer size for genome comparisons (the reference keeps it in a static, too) */
    private static int kmerSize = 21;

    private final String genomeId;
    private final String genomeName;

    public static void setKmerSize(int newSize) {
        kmerSize = newSize;
    }

    public static int getKmerSize() {
        return kmerSize;
    }

    /**
     * @throws Exception kept for source compatibility: the reference's constructor can throw checked
     * exceptions from the digest it computes (GenomeProcessor.java:112-114); this one does not
     */
    public GenomeKmers(Genome genome) throws Exception {
        this(GpuKmerEngine.shared(GpuKmerEngine.DNA, kmerSize), genome);
    }

    private GenomeKmers(GpuKmerEngine engine, Genome genome) {
        super(engine, engine.addGenome(contigSequences(genome)));
        this.genomeId = genome.getId();
        this.genomeName = genome.getName();
    }

    private static List<String> contigSequences(Genome genome) {
        List<String> retVal = new ArrayList<String>();
        for (Contig contig : genome.getContigs())
            retVal.add(contig.getSequence());
        return retVal;
    }

    public String getGenomeId() {
        return this.genomeId;
    }

    public String getGenomeName() {
        return this.genomeName;
    }
}
