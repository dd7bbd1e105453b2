package org.theseed.genome.distance;

import java.io.File;
import java.io.FileNotFoundException;
import java.io.IOException;
import java.io.PrintWriter;
import java.util.ArrayList;
import java.util.List;

import org.kohsuke.args4j.Argument;
import org.kohsuke.args4j.Option;
import org.slf4j.Logger;
import org.slf4j.LoggerFactory;
import org.theseed.basic.BaseReportProcessor;
import org.theseed.basic.ParseFailureException;
import org.theseed.genome.Genome;
import org.theseed.genome.iterator.GenomeSource;
import org.theseed.sequence.GenomeKmers;
import org.theseed.sequence.SequenceKmers;

/**
 * GPU replacement body for the "genomes" command (reference: GenomeProcessor.java).  Register it in App.java
 * in place of GenomeProcessor (case "genomes", App.java:52-54).
 *
 * Unchanged on purpose: options (-K/--kmerSize/--kmer default 21, -m/--maxDist default 0.9 -- validated and,
 * exactly as in the reference (:143-146), never applied --, -t/--type default DIR), validation messages, the
 * header and the three tab-separated columns in (source order x load order).
 * Changed: the base genomes become device-resident sets once; every query genome is compared with all of
 * them by ONE batched call (SequenceKmers.distances) instead of a parallel stream of per-pair probes
 * (:140), and its set is dropped afterwards.
 */
public class GpuGenomeProcessor extends BaseReportProcessor {

    protected static Logger log = LoggerFactory.getLogger(GpuGenomeProcessor.class);

    private List<GenomeKmers> mainKmers;

    @Option(name = "--kmerSize", aliases = { "-K", "--kmer" }, metaVar = "12", usage = "DNA kmer size")
    private int kmerSize;

    @Option(name = "--maxDist", aliases = { "-m", "--max", "--distance" }, metaVar = "0.75",
            usage = "maximum acceptable distance for a neighboring genome")
    private double maxDist;

    @Option(name = "--type", aliases = { "-t" }, usage = "genome source type")
    private GenomeSource.Type sourceType;

    @Argument(index = 0, metaVar = "gtoDir", required = true, usage = "base genome source")
    private File baseDir;

    @Argument(index = 1, metaVar = "gtoDir1 gtoDir2 ...", required = true, usage = "directory of input GTOs")
    private List<File> genomeDirs;

    @Override
    protected void setReporterDefaults() {
        this.kmerSize = 21;
        this.maxDist = 0.9;
        this.sourceType = GenomeSource.Type.DIR;
    }

    @Override
    protected void validateReporterParms() throws IOException, ParseFailureException {
        if (this.kmerSize < 4)
            throw new ParseFailureException("Kmer size cannot be less than 4.");
        GenomeKmers.setKmerSize(this.kmerSize);
        if (this.maxDist <= 0.0 || this.maxDist > 1.0)
            throw new ParseFailureException("Maximum distance must be > 0 and <= 1.");
        if (!this.baseDir.exists())
            throw new FileNotFoundException("Main genome source \"" + this.baseDir + "\" is not found.");
        for (File genomeDir : this.genomeDirs) {
            if (!genomeDir.exists())
                throw new FileNotFoundException("Genome source \"" + genomeDir + "\" is not found.");
        }
        try {
            GenomeSource baseGenomes = this.sourceType.create(this.baseDir);
            this.mainKmers = new ArrayList<GenomeKmers>(baseGenomes.size());
            log.info("Loading {} genomes from {}.", baseGenomes.size(), this.baseDir);
            for (Genome genome : baseGenomes)
                this.mainKmers.add(new GenomeKmers(genome));      // queued; built with the first query
        } catch (Exception e) {
            throw new ParseFailureException(e.toString());
        }
    }

    @Override
    protected void runReporter(PrintWriter writer) throws Exception {
        writer.println("genome1\tgenome2\tdistance");
        final SequenceKmers[] refs = this.mainKmers.toArray(new SequenceKmers[0]);
        final int nMain = refs.length;
        int compares = 0;
        for (File dir : this.genomeDirs) {
            log.info("Loading genome directory {}.", dir);
            GenomeSource genomes = this.sourceType.create(dir);
            for (Genome genome : genomes) {
                GenomeKmers kmers = new GenomeKmers(genome);
                double[] distances = SequenceKmers.distances(kmers, refs);    // one launch: query x all bases
                String genomeId = genome.getId();
                for (int i = 0; i < nMain; i++) {
                    writer.println(genomeId + "\t" + this.mainKmers.get(i).getGenomeId() + "\t" + distances[i]);
                    compares++;
                }
                kmers.getEngine().truncate(nMain);                            // drop the query's set
            }
        }
        log.info("{} comparisons output.", compares);
    }
}
