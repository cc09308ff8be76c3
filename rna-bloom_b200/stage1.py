"""rnabloom-gpu-stage1: RNA-Bloom's stage 1 (graph construction) on the GPU, writing exactly the files the unmodified JAR resumes from
(seam B2 of SURVEY.md section 8b; `RNABloom.java:6449-6452,7058,7134-7186`):

    <outdir>/<name>.graph{,.dbgbf,.dbgbf.desc,.cbf,.cbf.desc,.rpkbf,.rpkbf.desc}     graph/BloomFilterDeBruijnGraph.java:297-339
    <outdir>/<name>.readstats                                                       RNABloom.java:2669-2679  (min / Q1 / M / Q3 / max)
    <outdir>/DBG.DONE                                                               RNABloom.java:5818 (stage stamp)

The control loop is the reference's: sample the read lengths, populate the graph from the left reads (forward) and the right reads
(reverse-complemented under -revcomp-right when stranded), check `withinMaxFPR` (`RNABloom.java:1348-1353`: every filter's FPR <= 2 x
the requested FPR), and if it fails size the filters from the largest pop count (`getOptimalBloomFilterSizes` :1355-1385,
`BloomFilter.getExpectedSize` :196-199) and repopulate (:7142-7180).  Popcounts and FPRs come straight from the device
(`rb_filter_popcount`), inserts go through `rb_graph_add_reads_ascii` (segmentation by quality / ACGT runs on the GPU).

    python -m rnabloom_b200.stage1 -left L.fq[.gz] -right R.fq[.gz] -revcomp-right -stranded -k 25 -outdir out -name rnabloom

Host plumbing only (argument names follow the reference's CLI); the product is the library.
"""
import argparse
import gzip
import os
import sys

NUM_BITS_1GB = 8 * 1073741824.0      # RNABloom.java NUM_BITS_1GB
NUM_BYTES_1GB = 1073741824.0


def _open(path):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path)


def records(path):
    """(seq, qual or None) of a FASTQ / FASTA file (plain or gzip), one record at a time."""
    with _open(path) as fh:
        first = fh.read(1)
        if not first:
            return
        if first == "@":
            rest = fh.readline()
            while True:
                seq = fh.readline().rstrip("\n")
                fh.readline()
                qual = fh.readline().rstrip("\n")
                if not qual and not seq:
                    return
                yield seq, qual
                rest = fh.readline()
                if not rest:
                    return
        elif first == ">":
            fh.readline()
            seq = []
            for line in fh:
                if line.startswith(">"):
                    yield "".join(seq), None
                    seq = []
                else:
                    seq.append(line.strip())
            if seq:
                yield "".join(seq), None
        else:
            raise ValueError("%s is neither FASTQ nor FASTA" % path)


def quartiles(lengths):
    """util/Common.java:134-163 (integer arithmetic, sorted array)."""
    arr = sorted(lengths)
    n = len(arr)
    half, q1i = n // 2, n // 4
    q3i = half + q1i
    med = (arr[half - 1] + arr[half]) // 2 if n % 2 == 0 else arr[half]
    if n % 4 == 0:
        q1, q3 = (arr[q1i - 1] + arr[q1i]) // 2, (arr[q3i - 1] + arr[q3i]) // 2
    else:
        q1, q3 = arr[q1i], arr[q3i]
    return arr[0], q1, med, q3, arr[-1]


def read_length_quartiles(paths, k, sample):
    """RNABloom.java:1034-1099: per-file quartiles of the first `sample` reads with length >= k, Q1 / M / Q3 weighted by file size."""
    per_file, sizes = [], []
    for p in paths:
        lens = []
        for seq, _ in records(p):
            if len(seq) >= k:
                lens.append(len(seq))
                if len(lens) >= sample:
                    break
        if not lens:
            raise ValueError("Cannot determine read length from `%s`" % p)
        per_file.append(quartiles(lens))
        sizes.append(os.path.getsize(p))
    total = float(sum(sizes))
    w = [s / total for s in sizes]
    rint = lambda x: int(round(x / 2.0) * 2) if abs(x - int(x)) == 0.5 else int(round(x))   # Math.rint: ties to even
    return (min(q[0] for q in per_file), rint(sum(wi * q[1] for wi, q in zip(w, per_file))), rint(sum(wi * q[2] for wi, q in zip(w, per_file))),
            rint(sum(wi * q[3] for wi, q in zip(w, per_file))), max(q[4] for q in per_file))


def write_quartiles(q, path):
    """RNABloom.java:2669-2679."""
    with open(path, "w") as fh:
        fh.write("min:%d\nQ1:%d\nM:%d\nQ3:%d\nmax:%d\n" % q)


class Stage1:
    def __init__(self, ctx, k, stranded, dbgbf_bits, cbf_bytes, pkbf_bits, dbgbf_num_hash=2, cbf_num_hash=2, pkbf_num_hash=2, min_base_qual=3,
                 chunk_reads=1_000_000):
        self.ctx, self.k, self.stranded = ctx, k, stranded
        self.sizes = [dbgbf_bits, cbf_bytes, pkbf_bits]
        self.num_hash = (dbgbf_num_hash, cbf_num_hash, pkbf_num_hash)
        self.min_base_qual, self.chunk_reads = min_base_qual, chunk_reads
        self.graph = None
        self.read_pair_distance = -1

    def _new_graph(self, use_pairs):
        import rnabloom_b200 as rb
        if self.graph is not None:
            self.graph.destroy()                                          # destroyAllBf (:7151)
        hd, hc, hp = self.num_hash
        self.graph = rb.BloomFilterDeBruijnGraph(self.ctx, self.sizes[0], self.sizes[1], self.sizes[2], hd, hc, hp, self.k, self.stranded, use_pairs)
        if use_pairs:
            self.graph.setPairedKmerDistances(self.read_pair_distance, -1)

    def _add_file(self, path, revcomp, store_pairs):
        import rnabloom_b200 as rb
        flags = (rb.REVCOMP if revcomp else 0) | (rb.STORE_READ_PAIRS if store_pairs else 0)
        seqs, quals, n = [], [], 0
        fastq = None
        for seq, qual in records(path):
            if fastq is None:
                fastq = qual is not None
            seqs.append(seq)
            if fastq:
                quals.append(qual)
            if len(seqs) >= self.chunk_reads:
                n += self.graph.addReadsAscii(seqs, quals if fastq else None, self.min_base_qual, flags)
                seqs, quals = [], []
        if seqs:
            n += self.graph.addReadsAscii(seqs, quals if fastq else None, self.min_base_qual, flags)
        return n

    def populate(self, left, right, revcomp_right, store_pairs):
        """populateGraph2 for short reads (RNABloom.java:1159-1345): left reads forward, right reads reverse-complemented when asked."""
        n = 0
        for p in left:
            n += self._add_file(p, False, store_pairs)
        for p in right:
            n += self._add_file(p, bool(revcomp_right), store_pairs)
        return n

    def fprs(self):
        g = self.graph
        out = {"dbgbf": g.getDbgbfFPR(), "cbf": g.getCbfFPR()}
        if g.getRpkbf() is not None:
            out["rpkbf"] = g.getRpkbf().getFPR()
        return out

    def within_max_fpr(self, fpr):
        return all(v <= 2.0 * fpr for v in self.fprs().values())          # RNABloom.java:1348-1353

    def optimal_sizes(self, max_fpr):
        """getOptimalBloomFilterSizes (:1355-1385): every filter sized for the LARGEST pop count among them."""
        import rnabloom_b200 as rb
        g = self.graph
        pops = [g.getDbgbf().getPopCount(), g.getCbf().getPopCount()]
        if g.getRpkbf() is not None:
            pops.append(g.getRpkbf().getPopCount())
        m = max(pops)
        return [rb.BloomFilter.getExpectedSize(m, max_fpr, h) for h in self.num_hash]

    def histogram(self, paths, sample_bits=11, table_slots=1 << 24):
        """The k-mer multiplicity histogram the reference gets from `ntcard` (RNABloom.java:5745-5768), by hash sampling on the GPU."""
        import rnabloom_b200 as rb
        h = rb.KmerHistogram(self.ctx, self.k, self.stranded, sample_bits, table_slots)
        for p in paths:
            seqs = []
            for seq, _ in records(p):
                seqs.append(seq)
                if len(seqs) >= self.chunk_reads:
                    h.addReads(rb.pack_reads(seqs))
                    seqs = []
            if seqs:
                h.addReads(rb.pack_reads(seqs))
        h.finish()
        return h

    def sizes_from_histogram(self, hist, max_fpr):
        """RNABloom.java:6963,6986-7010: dbgbf / pkbf for F0 k-mers, cbf for the non-singletons."""
        import rnabloom_b200 as rb
        unexpected = (hist.numKmers > 0 and hist.numUniqueKmers < 1) or hist.numKmers < hist.numUniqueKmers
        exp = max(hist.numKmers, hist.numUniqueKmers) if unexpected else hist.numUniqueKmers
        hd, hc, hp = self.num_hash
        singletons = hist.getNumSingletons()
        cbf_n = exp if (exp == singletons or unexpected) else exp - singletons
        return [rb.BloomFilter.getExpectedSize(exp, max_fpr, hd), rb.BloomFilter.getExpectedSize(cbf_n, max_fpr, hc), rb.BloomFilter.getExpectedSize(exp, max_fpr, hp)]

    def run(self, left, right, revcomp_right, outdir, name, max_fpr=0.01, sample=1000, save=True, use_pairs=True, ntcard=False):
        os.makedirs(outdir, exist_ok=True)
        if ntcard:   # size the filters from the histogram and leave it where the JAR looks for ntcard's output (:5747-5750, :6934)
            hist = self.histogram(list(left) + list(right))
            hist.write(os.path.join(outdir, "%s_k%d.hist" % (name, self.k)))
            self.sizes = self.sizes_from_histogram(hist, max_fpr)
            hist.destroy()
        q = read_length_quartiles(list(left) + list(right), self.k, sample)
        write_quartiles(q, os.path.join(outdir, name + ".readstats"))
        # setReadLengthBasedParams (RNABloom.java:1017-1031): readPairedKmerDistance = max(1, Q1 - k - minNumKmerPairs(10))
        self.read_pair_distance = max(1, q[1] - self.k - 10)
        self._new_graph(use_pairs)
        n = self.populate(left, right, revcomp_right, use_pairs)
        report = {"kmers": n, "sizes": list(self.sizes), "fpr": self.fprs(), "resized": False}
        if not self.within_max_fpr(max_fpr):                              # :7142-7180
            self.sizes = self.optimal_sizes(max_fpr)
            self._new_graph(use_pairs)
            n = self.populate(left, right, revcomp_right, use_pairs)
            report.update({"kmers": n, "sizes": list(self.sizes), "fpr": self.fprs(), "resized": True})
        if save:
            self.graph.save(os.path.join(outdir, name + ".graph"))       # saveGraph (:7182-7184)
            open(os.path.join(outdir, "DBG.DONE"), "w").close()           # touch(dbgDoneStamp)
        report["readstats"] = q
        return report


def main(argv=None):
    ap = argparse.ArgumentParser(prog="rnabloom-gpu-stage1", description=__doc__.split("\n\n")[0])
    ap.add_argument("-left", nargs="+", default=[])
    ap.add_argument("-right", nargs="+", default=[])
    ap.add_argument("-revcomp-right", action="store_true", dest="revcomp_right")
    ap.add_argument("-stranded", action="store_true")
    ap.add_argument("-k", type=int, default=25)
    ap.add_argument("-outdir", default=os.path.join(os.getcwd(), "rnabloom_assembly"))
    ap.add_argument("-name", default="rnabloom")
    ap.add_argument("-fpr", type=float, default=0.01)
    ap.add_argument("-q", type=int, default=3, help="minimum base quality (PHRED33)")
    ap.add_argument("-dbgbf-gb", type=float, default=0.5, dest="dbgbf_gb")
    ap.add_argument("-cbf-gb", type=float, default=0.5, dest="cbf_gb")
    ap.add_argument("-pkbf-gb", type=float, default=0.25, dest="pkbf_gb")
    ap.add_argument("-hash", type=int, default=2, help="number of hash functions of every filter")
    ap.add_argument("-ntcard", action="store_true", help="size the filters from a k-mer histogram computed on the GPU (what RNA-Bloom runs ntcard for)")
    ap.add_argument("-device", type=int, default=0)
    a = ap.parse_args(argv)
    import rnabloom_b200 as rb
    ctx = rb.Context(a.device)
    s1 = Stage1(ctx, a.k, a.stranded, int(a.dbgbf_gb * NUM_BITS_1GB), int(a.cbf_gb * NUM_BYTES_1GB), int(a.pkbf_gb * NUM_BITS_1GB), a.hash, a.hash, a.hash, a.q)
    rep = s1.run(a.left, a.right, a.revcomp_right, a.outdir, a.name, a.fpr, ntcard=a.ntcard)
    print("stage 1 on the GPU: %d k-mers, sizes %s%s, FPR %s -> %s" % (rep["kmers"], rep["sizes"], " (resized)" if rep["resized"] else "", rep["fpr"],
                                                                      os.path.join(a.outdir, a.name + ".graph")))
    s1.graph.destroy()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
