package org.theseed.genome.distance;

import java.io.File;
import java.io.FileNotFoundException;
import java.io.IOException;
import java.io.PrintWriter;
import java.util.ArrayList;
import java.util.List;

import org.kohsuke.args4j.Option;
import org.slf4j.Logger;
import org.slf4j.LoggerFactory;
import org.theseed.basic.BaseReportProcessor;
import org.theseed.basic.ParseFailureException;
import org.theseed.sequence.FastaInputStream;
import org.theseed.sequence.KmerType;
import org.theseed.sequence.Sequence;
import org.theseed.sequence.gpu.GpuKmerEngine;

/**
 * GPU replacement body for the "fastaDist" command (reference: FastaDistanceProcessor.java).  Register it in
 * App.java in place of FastaDistanceProcessor (case "fastaDist", App.java:97-99).
 *
 * Unchanged on purpose: the options and their aliases (-i/--input, -K/--kSize/--kmerSize, -b/--batch, --type,
 * plus the framework's -o -h -v), the defaults (K 0 = type default, batch 20, DNA), the two validation
 * messages, the header and the five tab-separated columns with "" + distance number formatting.
 * Changed: the batch cache, the parallel row stream and computePairs (reference :141-194) collapse into one
 * engine build and block-wise distance calls; output order is deterministic list order (the reference
 * interleaves rows arbitrarily under its lock).  -b is still validated but only bounds the report block size.
 */
public class GpuFastaDistanceProcessor extends BaseReportProcessor {

    protected static Logger log = LoggerFactory.getLogger(GpuFastaDistanceProcessor.class);
    /** pairs fetched from the engine per call */
    private static final int BLOCK = 1 << 22;

    private List<Sequence> sequences;

    @Option(name = "--input", aliases = { "-i" }, usage = "input FASTA file (if not STDIN)")
    private File inFile;

    @Option(name = "--kSize", aliases = { "--kmerSize", "-K" }, usage = "kmer size to use; 0 for sequence type default")
    private int kmerSize;

    @Option(name = "--batch", aliases = { "-b" }, usage = "batch size for kmer cache and parallelism")
    private int batchSize;

    @Option(name = "--type", usage = "input sequence type")
    private KmerType seqType;

    @Override
    protected void setReporterDefaults() {
        this.inFile = null;
        this.kmerSize = 0;
        this.batchSize = 20;
        this.seqType = KmerType.DNA;
    }

    @Override
    protected void validateReporterParms() throws IOException, ParseFailureException {
        if (this.kmerSize == 0)
            this.kmerSize = this.seqType.getKmerSize();
        if (this.kmerSize < 2)
            throw new ParseFailureException("Kmer size must be at least 2.");
        if (this.batchSize < 1)
            throw new ParseFailureException("Batch size must be at least 1.");
        if (this.inFile != null && !this.inFile.canRead())
            throw new FileNotFoundException("Input file " + this.inFile + " is not found or unreadable.");
        try (FastaInputStream inStream = (this.inFile == null ? new FastaInputStream(System.in)
                : new FastaInputStream(this.inFile))) {
            this.sequences = new ArrayList<Sequence>();
            for (Sequence seq : inStream)
                this.sequences.add(seq);
            log.info("{} sequences read from input.", this.sequences.size());
        }
    }

    @Override
    protected void runReporter(PrintWriter writer) throws Exception {
        writer.println("seq1\tname1\tseq2\tname2\tdistance");
        final int n = this.sequences.size();
        long pairCount = 0;
        try (GpuKmerEngine engine = new GpuKmerEngine(this.seqType.getAlphabet(), this.kmerSize, 0)) {
            for (Sequence seq : this.sequences)
                engine.add(seq.getSequence());
            engine.build();                                   // kernels 1-3 for every sequence at once
            log.info("{} kmer sets built. Computing distances.", n);
            final long total = (long) n * (n - 1) / 2;
            int i = 0, j = 1;                                 // pair of linear index pairCount
            while (pairCount < total) {
                int count = (int) Math.min((long) BLOCK, total - pairCount);
                double[] dist = engine.allVsAllRange(pairCount, count);   // kernels 4-5
                for (int t = 0; t < count; t++) {
                    Sequence seq = this.sequences.get(i), seq2 = this.sequences.get(j);
                    writer.println(seq.getLabel() + "\t" + seq.getComment() + "\t" + seq2.getLabel() + "\t"
                            + seq2.getComment() + "\t" + dist[t]);
                    if (++j == n) {
                        i++;
                        j = i + 1;
                    }
                }
                pairCount += count;
            }
        }
        log.info("{} pairs computed.", pairCount);
    }
}
