"""Second, independent restatement of the hot path in pure Python (small cases only).

TEST INFRASTRUCTURE ONLY.  Written from the Java sources, not from rnabloom_oracle.c, and on purpose
in the *other* formulation the reference offers (seed table + Long.rotateLeft/Right, NTHash.java:188-300)
so that an agreement between the two restatements is meaningful.  tests/golden/make_golden.py freezes its
outputs into tests/golden/kat.json.  PARITY UNPINNED upstream (no JVM here): see DESIGN.md "Oracle".
Citations are relative to /root/reference/src/rnabloom/.
"""
M64 = (1 << 64) - 1

SEED = {"A": 0x3C8BFBB395C60474, "C": 0x3193C18562A02B4C, "G": 0x20323ED082572324, "T": 0x295549F54BE24456}
SEED["U"] = SEED["T"]
for _c in "ACGTU":
    SEED[_c.lower()] = SEED[_c]
COMPLEMENT = {"A": "T", "C": "G", "G": "C", "T": "A", "U": "A"}
MULTI_SEED = 0x90B45D39FB6DA1FA


def seed_of(ch):
    """seedTab[ch] (NTHash.java:135-168); rows 1,3,4,5,7 alias T,G,A,A,C for the complement trick."""
    o = ord(ch)
    if o in (1, 3, 4, 5, 7):
        return {1: SEED["T"], 3: SEED["G"], 4: SEED["A"], 5: SEED["A"], 7: SEED["C"]}[o]
    return SEED.get(ch, 0)


def seed_of_complement(ch):
    """seedTab[ch & cpOff] (NTHash.java:30,213)."""
    return seed_of(chr(ord(ch) & 7))


def rotl(v, s):
    s %= 64
    return ((v << s) | (v >> (64 - s))) & M64 if s else v


def rotr(v, s):
    return rotl(v, (64 - (s % 64)) % 64)


def signed(v):
    v &= M64
    return v - (1 << 64) if v >> 63 else v


def fhval(kmer):  # NTHash.java:188-193
    k = len(kmer)
    h = 0
    for i, ch in enumerate(kmer):
        h ^= rotl(seed_of(ch), k - 1 - i)
    return h


def rhval(kmer):  # NTHash.java:208-213
    h = 0
    for i, ch in enumerate(kmer):
        h ^= rotl(seed_of_complement(ch), i)
    return h


def canonical(f, r):  # NTHash.java:263  (signed compare)
    return r if signed(r) < signed(f) else f


def multi(base, k, m):  # NTHash.java:518-527
    out = [base & M64]
    for i in range(1, m):
        t = (base * ((i ^ ((k * MULTI_SEED) & M64)) & M64)) & M64
        t ^= t >> 27
        out.append(t)
    return out


def combine(a, b):  # HashFunction.java:260-266 ; int literal -1640531527 sign-extends
    a &= M64
    b &= M64
    lit = (-1640531527) & M64
    return a ^ ((b + lit + ((a << 6) & M64) + (b >> 2)) & M64)


def index(h, size):  # BloomFilter.java:108-111
    return ((h & M64) >> 1) % size


def kmer_bases(seq, k, mode, start=0, end=None):
    """(f, r, base) for every k-mer of seq[start:end]; mode 0 fwd, 1 rc, 2 canonical.  Direct (non-rolling)."""
    end = len(seq) if end is None else end
    out = []
    for p in range(start, end - k + 1):
        km = seq[p:p + k]
        f, r = fhval(km), rhval(km)
        b = f if mode == 0 else r if mode == 1 else canonical(f, r)
        out.append((f, r, b))
    return out


def pair_bases(seq, k, d, mode, start=0, end=None):  # Paired*NTHashIterator.java
    hs = kmer_bases(seq, k, mode, start, end)
    out = []
    for i in range(0, len(hs) - d):
        (fL, rL, _), (fR, rR, _) = hs[i], hs[i + d]
        if mode == 0:
            p = combine(fL, fR)
        elif mode == 1:
            p = combine(rR, rL)
        else:
            p1, p2 = combine(fL, fR), combine(rR, rL)
            p = p1 if signed(p1) <= signed(p2) else p2
        out.append(p)
    return out


# ---- MiniFloat (util/MiniFloat.java:27-45), deterministic part only (b < 16) ---------------------
def minifloat_to_float(b):
    if b <= 7:
        return float(b)
    return float(((b & 7) | 8) * 2 ** ((b >> 3) - 1))


def minifloat_increment_det(b):
    if b <= 15:
        return b + 1
    raise ValueError("probabilistic range; keep fixtures at multiplicity <= 17")


class PyBloom:  # BloomFilter.java + UnsafeBitBuffer.java
    def __init__(self, size, num_hash):
        self.size, self.h = size, num_hash
        self.bytes = bytearray((size + 7) // 8)

    def lookup_then_add(self, hv):
        found = True
        for x in hv[: self.h]:
            i = index(x, self.size)
            was = bool(self.bytes[i // 8] & (1 << (i % 8)))
            self.bytes[i // 8] |= 1 << (i % 8)
            found = was and found
        return found

    def add(self, hv):
        for x in hv[: self.h]:
            i = index(x, self.size)
            self.bytes[i // 8] |= 1 << (i % 8)

    def lookup(self, hv):
        return all(self.bytes[index(x, self.size) // 8] & (1 << (index(x, self.size) % 8)) for x in hv[: self.h])


class PyCounting:  # CountingBloomFilter.java:170-251
    def __init__(self, size, num_hash):
        self.size, self.h = size, num_hash
        self.bytes = bytearray(size)

    def increment(self, hv):
        idx = [index(x, self.size) for x in hv[: self.h]]
        m = min(self.bytes[i] for i in idx)
        u = minifloat_increment_det(m)
        for i in idx:
            if self.bytes[i] == m:
                self.bytes[i] = u

    def get_count(self, hv):
        return minifloat_to_float(min(self.bytes[index(x, self.size)] for x in hv[: self.h]))


class PyGraph:  # BloomFilterDeBruijnGraph.java:75-104,405-412,562-570
    def __init__(self, dbg_bits, cbf_bytes, hd, hc, k, stranded):
        self.k, self.stranded, self.hmax = k, stranded, max(hd, hc)
        self.dbgbf, self.cbf = PyBloom(dbg_bits, hd), PyCounting(cbf_bytes, hc)

    def add_seq(self, seq, revcomp=False):
        mode = 2 if not self.stranded else (1 if revcomp else 0)
        for _, _, b in kmer_bases(seq, self.k, mode):
            hv = multi(b, self.k, self.hmax)
            if self.dbgbf.lookup_then_add(hv):
                self.cbf.increment(hv)

    def counts(self, seq):
        mode = 0 if self.stranded else 2
        out = []
        for _, _, b in kmer_bases(seq, self.k, mode):
            hv = multi(b, self.k, self.hmax)
            out.append(self.cbf.get_count(hv) + 1 if self.dbgbf.lookup(hv) else 0.0)
        return out
