#!/usr/bin/env python
"""Freeze known-answer vectors from the pure-Python restatement (oracle/pyref.py) into kat.json.

Run from the repo root:  python tests/golden/make_golden.py
The reference has no golden vectors for this path and no JVM exists here (PARITY UNPINNED upstream); these
vectors pin the C oracle and the CUDA path to the second restatement and to SURVEY.md section 4's table.
"""
import json
import os
import random
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as P  # noqa: E402

H = lambda v: "0x%016x" % (v & P.M64)  # noqa: E731
rng = random.Random(20261017)


def rand_seq(n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def main():
    kat = {}
    # --- SURVEY.md section 4 scratch table (copied by hand from the survey, not recomputed) ----------------
    kat["survey"] = {
        "ntp64_ACGTx4_k16": "0xc62d19dfe0703f82",
        "k25_GATTACA": {"seq": "GATTACAGATTACAGATTACAGATT", "f": "0x0fd07f35ad194b4f", "r": "0x06cf853e559e9eca",
                        "multi3": ["0x06cf853e559e9eca", "0x573a7d41d7b8e108", "0x42cbed986966de16"],
                        "idx_2p33": [5013196645, 3957092484, 884174603],
                        "idx_8589934583": [5270289196, 7249857108, 3405661601]},
        "k35_GATTACA": {"seq": "GATTACAGATTACAGATTACAGATTACAGATTACA", "f": "0x6fba69ec61acc109", "r": "0xb40fd5470336d37f"},
        "k17_GATTACA": {"seq": "GATTACAGATTACAGAT", "f": "0x9af50c20b5bc5897", "r": "0x980ca43422437f61",
                        "multi3": ["0x980ca43422437f61", "0x348e51eef6077db7", "0x6c6865418c0e6b18"]},
        "combine_k25_pos0_pos10": {"seq": "GATTACA" * 6, "value": "0x8b36256e0925d5f4"},
    }
    # --- hashes ------------------------------------------------------------------------------------------
    seqs = ["GATTACA" * 30, rand_seq(150), rand_seq(150, "ACGTacgtU"), rand_seq(97) + "N" + rand_seq(120), "A" * 140]
    kat["hashes"] = []
    for s in seqs:
        for k in (17, 25, 35, 64, 65, 100):
            hs = P.kmer_bases(s, k, 2)
            kat["hashes"].append({"seq": s, "k": k, "f": [H(f) for f, _, _ in hs], "r": [H(r) for _, r, _ in hs],
                                  "canon": [H(b) for _, _, b in hs]})
    kat["multi"] = []
    for _ in range(24):
        b, k, m = rng.getrandbits(64), rng.choice([17, 25, 35, 64, 65, 100]), rng.choice([1, 2, 3, 4, 7])
        kat["multi"].append({"base": H(b), "k": k, "m": m, "hv": [H(x) for x in P.multi(b, k, m)]})
    kat["index"] = []
    sizes = [1, 2, 3, 1021, 2 ** 33, 2 ** 36, 8589934583, 68719476735, 2 ** 62 + 57, 2 ** 63 - 1, 2 ** 63 - 25, 10 ** 12 + 39]
    for size in sizes:
        for _ in range(6):
            h = rng.getrandbits(64)
            kat["index"].append({"h": H(h), "size": size, "idx": P.index(h, size)})
        for h in (0, 1, P.M64, 1 << 63, (size << 1) & P.M64, ((size << 1) - 1) & P.M64, ((size << 1) + 2) & P.M64):
            kat["index"].append({"h": H(h), "size": size, "idx": P.index(h, size)})
    kat["combine"] = []
    for _ in range(16):
        a, b = rng.getrandbits(64), rng.getrandbits(64)
        kat["combine"].append({"a": H(a), "b": H(b), "v": H(P.combine(a, b))})
    kat["pairs"] = []
    for s in seqs[:3]:
        for k, d in ((25, 10), (35, 105), (17, 1), (25, 125)):
            for mode in (0, 1, 2):
                kat["pairs"].append({"seq": s, "k": k, "d": d, "mode": mode, "p": [H(x) for x in P.pair_bases(s, k, d, mode)]})
    kat["minifloat_to_float"] = [P.minifloat_to_float(b) for b in range(128)]
    # --- segmentation (java.util.regex == python re for these classes) --------------------------------------
    phred = "!\"#$%&'()*+,-./0123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVWXYZ[\\]^_`abcdefghijklmnopqrstuvwxyz{|}~"
    kat["segments"] = []
    for _ in range(40):
        n = rng.choice([10, 24, 25, 26, 60, 150, 151])
        k = rng.choice([17, 25])
        mq = rng.choice([0, 3, 20])
        s = "".join(rng.choice("ACGTACGTACGTACGTacgtuUN.") if rng.random() < 0.12 else rng.choice("ACGT") for _ in range(n))
        q = "".join(chr(33 + (rng.choice([0, 1, 2, 5, 19]) if rng.random() < 0.06 else rng.randint(20, 41))) for _ in range(n))
        segs = []
        if n >= k:
            qp = re.compile("[" + re.escape(phred[mq:]) + "]{%d,}" % k)
            sp = re.compile("[ACGTU]{%d,}" % k, re.IGNORECASE)
            for mqm in qp.finditer(q):
                pos = mqm.start()
                while True:
                    m = sp.search(s, pos, mqm.end())
                    if not m:
                        break
                    segs.append([m.start(), m.end()])
                    pos = m.end()
        fasta = [[m.start(), m.end()] for m in re.finditer("[ACGTU]{%d,}" % k, s, re.IGNORECASE)] if n >= k else []
        kat["segments"].append({"seq": s, "qual": q, "k": k, "min_qual": mq, "fastq": segs, "fasta": fasta})
    # --- small graphs: sequential BloomFilterDeBruijnGraph.add then getCount ---------------------------------
    kat["graphs"] = []
    for stranded, k, hd, hc, dbg_bits, cbf_bytes in ((True, 25, 3, 3, 4099, 1021), (False, 25, 2, 3, 8191, 2048),
                                                   (False, 17, 3, 2, 65536, 4096), (True, 35, 1, 1, 1000, 333)):
        genome = rand_seq(400)
        reads = []
        for _ in range(30):
            p = rng.randrange(0, len(genome) - 80)
            reads.append(genome[p:p + 80])
        g = P.PyGraph(dbg_bits, cbf_bytes, hd, hc, k, stranded)
        for i, r in enumerate(reads):
            g.add_seq(r, revcomp=(stranded and i % 2 == 1))
        if max(g.cbf.bytes) > 16:
            raise SystemExit("fixture reached the probabilistic MiniFloat range")
        kat["graphs"].append({"stranded": stranded, "k": k, "hd": hd, "hc": hc, "dbg_bits": dbg_bits, "cbf_bytes": cbf_bytes,
                              "reads": reads, "dbgbf": bytes(g.dbgbf.bytes).hex(), "cbf": bytes(g.cbf.bytes).hex(),
                              "query": genome[:120], "counts": g.counts(genome[:120])})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
    with open(out, "w") as fh:
        json.dump(kat, fh, indent=0, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
