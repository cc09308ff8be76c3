import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rnabloom_b200 as rb
from oracle.binding import Oracle, OracleGraph
orc = Oracle()
ctx = rb.Context(0)

def full(n_reads, dbg_bits, cbf_bytes, genome):
    L, stride, k = 150, 160, 25
    nk = n_reads * (L - k + 1)
    g = rb.BloomFilterDeBruijnGraph(ctx, dbg_bits, cbf_bytes, 64, 3, 3, 1, k, False, False)
    packed = ctx.dev_alloc(n_reads * stride // 4 + 64)
    counts_dev = ctx.dev_alloc(nk * 4)
    ctx.synth_reads_dev(77, genome, 0, n_reads, L, 2000, stride, packed)
    g.addReadsDev(packed, n_reads, L, stride)
    c1 = np.zeros(nk, dtype=np.float32)
    g.getKmersDev(packed, n_reads, L, stride, counts_dev); ctx.sync(); ctx.d2h(c1, counts_dev)
    g.addReadsDev(packed, n_reads, L, stride)
    c2 = np.zeros(nk, dtype=np.float32)
    g.getKmersDev(packed, n_reads, L, stride, counts_dev); ctx.sync(); ctx.d2h(c2, counts_dev)
    small = c1 <= 8
    bad = small & (c2 != 2 * c1)
    print("full: n=%d nk=%d mismatches=%d (%.2e)" % (n_reads, nk, bad.sum(), bad.mean()))
    if bad.any():
        idx = np.nonzero(bad)[0][:20]
        print(" c1", c1[idx]); print(" c2", c2[idx])
        d = (c2 - 2 * c1)[bad]
        print(" diff hist", np.unique(d, return_counts=True))
    # compare with the oracle on a small slice
    if n_reads <= 20000:
        h = np.zeros(n_reads * stride // 32, dtype=np.uint64); ctx.d2h(h, packed)
        reads = orc.synth_reads(77, genome, 0, n_reads, L, 2000)
        og = OracleGraph(orc, dbg_bits, cbf_bytes, 64, 3, 3, 1, k, False, False)
        for _ in range(2):
            for r in reads: og.add_read(bytes(r))
        dd = np.nonzero(g.getCbf().download() != og.cbf())[0]
        print(" vs oracle: cbf diffs", len(dd), "dbg equal", (g.getDbgbf().download() == og.dbgbf()).all())
        if len(dd):
            print("  gpu", g.getCbf().download()[dd[:20]], "orc", og.cbf()[dd[:20]])
        og.close()
    ctx.dev_free(packed); ctx.dev_free(counts_dev); g.destroy()

full(20000, 1 << 32, 1 << 30, 1_000_000)
full(400000, 1 << 34, 1 << 31, 20_000_000)

def loaded():
    reads = orc.synth_reads(9, 20000, 0, 700, 150, 5000)
    seqs = [bytes(r) for r in reads]
    dbg_bits, cbf_bytes = 600_011, 150_001
    g = rb.BloomFilterDeBruijnGraph(ctx, dbg_bits, cbf_bytes, 64, 3, 3, 1, 25, False, False)
    g.addReads(rb.pack_reads(seqs))
    got = g.getCbf().download().astype(np.int16)
    rng = np.random.default_rng(0); lo = hi = None
    for p in range(4):
        og = OracleGraph(orc, dbg_bits, cbf_bytes, 64, 3, 3, 1, 25, False, False)
        order = np.arange(len(seqs)) if p == 0 else rng.permutation(len(seqs))
        for i in order: og.add_read(seqs[i])
        if p == 0: print("loaded: dbg equal", (g.getDbgbf().download() == og.dbgbf()).all())
        c = og.cbf().astype(np.int16)
        if p == 0: c0 = c
        lo = c if lo is None else np.minimum(lo, c); hi = c if hi is None else np.maximum(hi, c)
        og.close()
    print(" outside env(+-1): %.4f  outside env(0): %.4f  env width>0: %.4f  sum gpu %d sum oracle0 %d lo %d hi %d max %d" % (
        ((got < lo - 1) | (got > hi + 1)).mean(), ((got < lo) | (got > hi)).mean(), (hi > lo).mean(), got.sum(), c0.sum(), lo.sum(), hi.sum(), got.max()))
    d = (got - c0); print(" diff vs order0 hist", np.unique(d, return_counts=True))
loaded()
