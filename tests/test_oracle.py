"""The C oracle against the golden vectors (pure-Python restatement + SURVEY scratch table) and against
algebraic identities of ntHash.  CPU only."""
import numpy as np
import pytest

from oracle import pyref as P
from oracle.binding import MODE_CANON, MODE_FWD, MODE_RC, OracleGraph

U = lambda a: [int(x) & P.M64 for x in a]  # noqa: E731
hx = lambda s: int(s, 16)  # noqa: E731


def test_survey_scratch_vectors(orc, kat):
    s = kat["survey"]
    assert orc.ntp64("ACGT" * 4, 16) & P.M64 == hx(s["ntp64_ACGTx4_k16"])
    assert orc.ntp64rc("ACGT" * 4, 16) & P.M64 == hx(s["ntp64_ACGTx4_k16"])
    for key, k in (("k25_GATTACA", 25), ("k35_GATTACA", 35), ("k17_GATTACA", 17)):
        e = s[key]
        f, r, b = orc.kmer_hashes(e["seq"], k, MODE_CANON)
        assert U(f)[0] == hx(e["f"]) and U(r)[0] == hx(e["r"])
        assert U(b)[0] == hx(e["r"])  # canonical = r in all three (k=35 only under a SIGNED compare)
        if "multi3" in e:
            assert U(orc.ntm64(b[0], k, 3)) == [hx(x) for x in e["multi3"]]
        if "idx_2p33" in e:
            hv = orc.ntm64(b[0], k, 3)
            assert [orc.index(x, 2 ** 33) for x in hv] == e["idx_2p33"]
            assert [orc.index(x, 8589934583) for x in hv] == e["idx_8589934583"]
    e = s["combine_k25_pos0_pos10"]
    f, _, _ = orc.kmer_hashes(e["seq"], 25, MODE_FWD)
    assert orc.combine(f[0], f[10]) & P.M64 == hx(e["value"])


def test_tables_are_rotations(orc):
    for c in list(b"ACGTUacgtuN") + [0, 1, 3, 4, 5, 7, 255]:
        for i in range(64):
            assert orc.lib.orc_mstab(c, i) == P.rotl(orc.lib.orc_seed(c), i)
    assert orc.lib.orc_seed(ord("N")) == 0
    for ch in "ACGTU":
        assert orc.lib.orc_seed(ord(ch) & 7) == P.SEED[P.COMPLEMENT[ch]]
        assert orc.lib.orc_seed(ord(ch.lower()) & 7) == P.SEED[P.COMPLEMENT[ch]]


def test_hashes_match_golden(orc, kat):
    for e in kat["hashes"]:
        f, r, b = orc.kmer_hashes(e["seq"], e["k"], MODE_CANON)
        assert U(f) == [hx(x) for x in e["f"]]
        assert U(r) == [hx(x) for x in e["r"]]
        assert U(b) == [hx(x) for x in e["canon"]]
        f2, _, b2 = orc.kmer_hashes(e["seq"], e["k"], MODE_FWD)
        assert U(f2) == U(f) and U(b2) == U(f)
        _, r3, b3 = orc.kmer_hashes(e["seq"], e["k"], MODE_RC)
        assert U(r3) == U(r) and U(b3) == U(r)


def test_roll_equals_direct_and_rc_symmetry(orc):
    rng = np.random.default_rng(7)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    for k in (17, 25, 35, 64, 65, 100):
        s = "".join(rng.choice(list("ACGT"), size=260))
        f, r, b = orc.kmer_hashes(s, k, MODE_CANON)
        rc = "".join(comp[c] for c in reversed(s))
        f2, r2, b2 = orc.kmer_hashes(rc, k, MODE_CANON)
        for p in range(len(s) - k + 1):
            assert f[p] == orc.ntp64(s, k, p)       # rolled == direct
            assert r[p] == orc.ntp64rc(s, k, p)
            q = len(s) - k - p
            assert r[p] == f2[q] and f[p] == r2[q]   # rev(s) == fwd(revcomp(s))
            assert b[p] == b2[q]                     # canonical is strand symmetric
        # a sub-range restarts the recurrence at `start`
        fs, _, _ = orc.kmer_hashes(s, k, MODE_FWD, 13, 13 + k + 40)
        assert list(fs) == list(f[13:13 + 41])


def test_multi_index_combine_golden(orc, kat):
    for e in kat["multi"]:
        assert U(orc.ntm64(P.signed(hx(e["base"])), e["k"], e["m"])) == [hx(x) for x in e["hv"]]
    for e in kat["index"]:
        assert orc.index(P.signed(hx(e["h"])), e["size"]) == e["idx"]
    for e in kat["combine"]:
        assert orc.combine(P.signed(hx(e["a"])), P.signed(hx(e["b"]))) & P.M64 == hx(e["v"])


def test_pairs_golden(orc, kat):
    for e in kat["pairs"]:
        _, _, p = orc.pair_hashes(e["seq"], e["k"], e["d"], e["mode"])
        assert U(p) == [hx(x) for x in e["p"]]


def test_minifloat(orc, kat):
    tab = kat["minifloat_to_float"]
    for b in range(128):
        assert orc.lib.orc_minifloat_to_float(b) == tab[b]
    assert [tab[i] for i in (7, 8, 15, 16, 17, 24, 127)] == [7.0, 8.0, 15.0, 16.0, 18.0, 32.0, 245760.0]
    for b in range(16):
        assert orc.lib.orc_minifloat_increment(b) == b + 1  # deterministic range (MiniFloat.java:31-38)
    assert orc.lib.orc_minifloat_increment(127) == 127
    orc.lib.orc_seed_rng(99)
    for b, p in ((16, 0.5), (24, 0.25), (40, 1 / 16)):
        hits = sum(orc.lib.orc_minifloat_increment(b) == b + 1 for _ in range(20000))
        assert abs(hits / 20000 - p) < 0.02


def test_segmentation_golden(orc, kat):
    for e in kat["segments"]:
        assert [list(x) for x in orc.segment(e["seq"], e["qual"], e["k"], e["min_qual"])] == e["fastq"]
        assert [list(x) for x in orc.segment(e["seq"], None, e["k"])] == e["fasta"]


def test_graph_golden(orc, kat):
    for e in kat["graphs"]:
        g = OracleGraph(orc, e["dbg_bits"], e["cbf_bytes"], 64, e["hd"], e["hc"], 1, e["k"], e["stranded"], False)
        for i, r in enumerate(e["reads"]):
            g.add_read(r, flags=1 if (e["stranded"] and i % 2 == 1) else 0)
        assert bytes(g.dbgbf()).hex() == e["dbgbf"]
        assert bytes(g.cbf()).hex() == e["cbf"]
        counts, _, _ = g.count_seq(e["query"])
        assert counts.tolist() == e["counts"]
        g.close()


def test_graph_policies_and_pairs(orc):
    rng = np.random.default_rng(3)
    s = "".join(rng.choice(list("ACGT"), size=300))
    k, d = 25, 10
    g = OracleGraph(orc, 1 << 20, 1 << 18, 1 << 18, 3, 3, 2, k, False, True)
    g.set_distances(d, 40)
    g.init_fpkbf(1 << 16, 2)
    n = g.add_read(s, flags=8 | 16)
    assert n == len(s) - k + 1
    counts, _, _ = g.count_seq(s)
    assert (counts == 1.0).all()           # first sighting => 1 (graph :562-570)
    g.add_read(s)
    g.add_read(s)
    counts, _, _ = g.count_seq(s)
    assert (counts == 3.0).all()
    before = g.cbf().copy()
    g.add_read(s, flags=4)                 # addDbgOnly leaves the counting filter alone
    assert (g.cbf() == before).all()
    g.add_read(s, flags=2)                 # addCountIfPresent: present with count > 0 => increment
    counts, _, _ = g.count_seq(s)
    assert (counts == 4.0).all()
    # pair filters hold exactly the pair hashes of the segment
    _, _, p = orc.pair_hashes(s, k, d, MODE_CANON)
    rp = orc.lib.orc_graph_rpkbf(g.g)
    assert all(orc.lib.orc_bf_lookup1(rp, int(x)) for x in p)
    assert orc.lib.orc_bf_popcount(rp) <= 2 * len(p)
    # a k-mer covering an invalid nucleotide counts 0 but is still emitted (HashFunction.java:55-85)
    t = s[:60] + "N" + s[61:140]
    counts, fh, _ = g.count_seq(t)
    assert len(counts) == len(t) - k + 1
    assert (counts[36:61] == 0).all() and (counts[:36] > 0).all() and (counts[61:] > 0).all()
    g.close()


def test_cascade_and_single_hash_overloads(orc):
    lib = orc.lib
    c = lib.orc_cascade_create(3000, 2, 25, 3)
    for _ in range(2):
        lib.orc_cascade_add1(c, 12345)
    assert lib.orc_bf_lookup1(lib.orc_cascade_level(c, 0), 12345) == 1
    assert lib.orc_bf_lookup1(lib.orc_cascade_level(c, 1), 12345) == 1
    assert lib.orc_cascade_lookup1(c, 12345) == 0
    lib.orc_cascade_add1(c, 12345)
    assert lib.orc_cascade_lookup1(c, 12345) == 1
    lib.orc_cascade_destroy(c)
    fpr = float(np.float32(0.01))  # Java widens the float argument to double before log()
    assert lib.orc_expected_size(10 ** 6, 0.01, 3) == int(np.ceil(10 ** 6 * (-3 / np.log(1 - np.exp(np.log(fpr) / 3)))))


def test_synth_reads_are_reproducible(orc):
    a = orc.synth_reads(1, 100000, 0, 50, 150, 5000)
    b = orc.synth_reads(1, 100000, 10, 10, 150, 5000)
    assert (a[10:20] == b).all()
    assert set(np.unique(a)) <= set(b"ACGT")


@pytest.mark.parametrize("threads", [1, 4])
def test_threaded_baseline_runs(orc, threads):
    reads = orc.synth_reads(5, 20000, 0, 400, 150, 5000)
    g = OracleGraph(orc, 1 << 24, 1 << 22, 64, 3, 3, 1, 25, False, False)
    km, _ = g.run_mt(reads, 0, False, threads)
    assert km == 400 * 126
    km, cs = g.run_mt(reads, 0, True, threads)
    assert km == 400 * 126 and cs >= km
    g.close()
