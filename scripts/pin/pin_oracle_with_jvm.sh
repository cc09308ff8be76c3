#!/bin/bash
# pin_oracle_with_jvm.sh -- pins the oracle AND the GPU path against the real reference, for anyone who has a JDK (the build image has
# none: `java`, `javac`, jni.h are absent, so DESIGN.md says "parity unpinned" until somebody runs this).
#
#   RNABLOOM_SRC=/path/to/RNA-Bloom/src  LIBS=/path/to/{commons-cli,jgrapht-core,smile-*}.jar  scripts/pin/pin_oracle_with_jvm.sh [workdir]
#
# 1. compiles the reference's bloom/ graph/ util/ classes and scripts/pin/PinOracle.java
# 2. JVM: sequences -> reference NTHash iterators -> graph.add / addReadSingleKmerPair -> graph.save  (jvm.graph*) + kat.tsv
# 3. GPU: the same sequences -> rb_graph_add_reads_ascii -> rb_graph_save                              (gpu.graph*)
# 4. cmp the .dbgbf / .rpkbf byte for byte (order-free filters: must be identical), the .cbf up to shared counters, the .desc grammar;
#    and re-checks the oracle (liboracle.so) against kat.tsv: every k-mer's hVals
set -euo pipefail
W=${1:-/tmp/rb_pin}; mkdir -p "$W/classes"
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
: "${RNABLOOM_SRC:?set RNABLOOM_SRC to the reference's src directory}"
CP="${LIBS:-}"
javac -nowarn -cp "$CP" -d "$W/classes" $(find "$RNABLOOM_SRC/rnabloom" -name '*.java') "$ROOT/scripts/pin/PinOracle.java"
python - "$W" <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from oracle.binding import Oracle
w = sys.argv[1]
reads = [bytes(r).decode() for r in Oracle().synth_reads(4242, 20000, 0, 400, 150, 5000)]
reads[3] = reads[3][:70] + "N" + reads[3][71:]
open(w + "/seqs.txt", "w").write("\n".join(reads) + "\n")
PY
K=25; DBG=$(( (1<<24) + 5 )); CBF=$(( (1<<22) + 1 )); PK=$(( (1<<20) + 7 )); H=3
java -cp "$CP:$W/classes" PinOracle "$W/seqs.txt" "$W" $K false $DBG $CBF $PK $H
python - "$W" $K $DBG $CBF $PK $H <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, rnabloom_b200 as rb
from oracle.binding import Oracle, MODE_CANON
w, k, dbg, cbf, pk, h = sys.argv[1], *map(int, sys.argv[2:])
seqs = open(w + "/seqs.txt").read().split()
ctx = rb.Context(0)
g = rb.BloomFilterDeBruijnGraph(ctx, dbg, cbf, pk, h, h, h, k, False, True)
g.setPairedKmerDistances(max(1, 150 - k - 10), -1)
g.addReadsAscii(seqs, None, 0, rb.STORE_READ_PAIRS)
g.save(w + "/gpu.graph")
orc = Oracle()
bad = 0
for line in open(w + "/kat.tsv"):
    _, kmer, kk, hv = line.rstrip("\n").split("\t")
    want = [int(x) for x in hv.strip("[]").split(", ")]
    base = orc.kmer_hashes(kmer, int(kk), MODE_CANON)[2][0]
    bad += list(orc.ntm64(base, int(kk), len(want))) != want
print("oracle vs JVM hash vectors: %d mismatches" % bad)
sys.exit(1 if bad else 0)
PY
for f in dbgbf rpkbf; do cmp "$W/jvm.graph.$f" "$W/gpu.graph.$f" && echo "$f: identical to the JVM's"; done
python - "$W" <<'PY'
import sys, numpy as np
w = sys.argv[1]
a, b = np.fromfile(w + "/jvm.graph.cbf", np.uint8), np.fromfile(w + "/gpu.graph.cbf", np.uint8)
print("cbf: %d of %d counters differ (shared counters are order dependent in the reference itself)" % ((a != b).sum(), len(a)))
for f in ("", ".dbgbf.desc", ".cbf.desc"):
    ja, gb = open(w + "/jvm.graph" + f).read().splitlines(), open(w + "/gpu.graph" + f).read().splitlines()
    same = [x for x, y in zip(ja, gb) if x == y or x.startswith("fpr:")]
    print("desc%s: %d / %d lines identical (fpr lines: Float.toString compared separately: %s vs %s)" % (f, len(same), len(ja), [x for x in ja if x.startswith("fpr")], [x for x in gb if x.startswith("fpr")]))
PY
