// PinOracle.java -- closes the "parity unpinned" gap for anyone with a JDK (none exists in the build image, so this file has never been
// compiled here).  It drives the REFERENCE's own classes -- rnabloom.graph.BloomFilterDeBruijnGraph, the NTHash iterators, graph.add /
// addReadSingleKmerPair / save -- over a plain text file of sequences, exactly the way FastaToGraphWorker does (RNABloom.java:677-716),
// and writes (1) the graph files (graph/BloomFilterDeBruijnGraph.java:297-339) and (2) known-answer vectors of the hash path.
// scripts/pin/pin_oracle_with_jvm.sh then builds the same graph with rb_graph_save and `cmp`s the files, and feeds the vectors to
// tests/test_oracle.py (RB_PIN_VECTORS=<file>).
//
//   javac -cp <RNA-Bloom classes or jar> -d out scripts/pin/PinOracle.java
//   java  -cp <...>:out PinOracle seqs.txt outdir k stranded dbgbfBits cbfBytes pkbfBits numHash
import java.io.*;
import java.nio.file.*;
import java.util.*;
import java.util.regex.*;
import rnabloom.bloom.hash.*;
import rnabloom.graph.BloomFilterDeBruijnGraph;

public class PinOracle {
    public static void main(String[] a) throws Exception {
        List<String> seqs = Files.readAllLines(Paths.get(a[0]));
        String outdir = a[1];
        int k = Integer.parseInt(a[2]);
        boolean stranded = Boolean.parseBoolean(a[3]);
        long dbgbfBits = Long.parseLong(a[4]), cbfBytes = Long.parseLong(a[5]), pkbfBits = Long.parseLong(a[6]);
        int h = Integer.parseInt(a[7]);
        new File(outdir).mkdirs();
        BloomFilterDeBruijnGraph graph = new BloomFilterDeBruijnGraph(dbgbfBits, cbfBytes, pkbfBits, h, h, h, k, stranded, true);
        int d = Math.max(1, 150 - k - 10);
        graph.setReadPairedKmerDistance(d);
        Pattern seqPattern = Pattern.compile("[ACGTUacgtu]{" + k + ",}");          // util/SeqUtils.java:1430-1438 getNucleotideCharsPattern
        NTHashIterator itr = graph.getHashIterator();
        PairedNTHashIterator pitr = graph.getPairedHashIterator(d);
        try (PrintWriter kat = new PrintWriter(new FileWriter(outdir + "/kat.tsv"))) {
            for (String seq : seqs) {
                Matcher m = seqPattern.matcher(seq);
                while (m.find()) {
                    itr.start(seq, m.start(), m.end());
                    while (itr.hasNext()) {
                        itr.next();
                        kat.println("kmer\t" + seq.substring(itr.getPos(), itr.getPos() + k) + "\t" + k + "\t" + Arrays.toString(itr.hVals));
                        graph.add(itr.hVals);
                    }
                    pitr.start(seq, m.start(), m.end());
                    while (pitr.hasNext()) {
                        pitr.next();
                        graph.addReadSingleKmerPair(pitr.hValsP);
                    }
                }
            }
        }
        graph.save(new File(outdir + "/jvm.graph"));
    }
}
