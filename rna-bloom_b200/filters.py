"""Host-side mirror of the reference's operator surface (same class and method names, argument meaning and error
behaviour) over the C-ABI.  The Java originals: bloom/BloomFilter.java, bloom/CountingBloomFilter.java,
graph/BloomFilterDeBruijnGraph.java (relative to /root/reference/src/rnabloom/).  Bulk methods take numpy arrays;
what was a per-k-mer Java call (`add(long[])`, `getCount(long)`) is the same call on an array of base hashes.
"""
import ctypes as C

import numpy as np

from . import binding as B
from .binding import RBError


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU: device memory, stream, claim table (rb_ctx)."""

    def __init__(self, device=0):
        self.L = B.lib()
        h = C.c_void_p()
        rc = self.L.rb_ctx_create(device, C.byref(h))
        if rc:
            raise RBError(rc, (self.L.rb_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def check(self, rc):
        if rc:
            raise RBError(rc, (self.L.rb_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.L.rb_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        self.check(self.L.rb_ctx_sync(self.h))

    def wait(self, ticket):
        """rb_ctx_wait: the results of the asynchronous call with this ticket (and of all earlier ones) are in host memory."""
        self.check(self.L.rb_ctx_wait(self.h, ticket))

    def set_stream(self, cuda_stream_ptr):
        self.check(self.L.rb_ctx_set_stream(self.h, cuda_stream_ptr))

    def set_rng_seed(self, seed):
        self.check(self.L.rb_ctx_set_rng_seed(self.h, seed))

    def set_subbatch_kmers(self, n):
        self.check(self.L.rb_ctx_set_subbatch_kmers(self.h, n))

    def kernel_launches(self):
        return self.L.rb_ctx_kernel_launches(self.h)

    def profile_enable(self, on=True):
        """Per-kernel CUDA-event timing of the read-level engines (measurement only)."""
        self.check(self.L.rb_ctx_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        """{kernel name: (summed ms, launches)} since the last read."""
        names = C.create_string_buffer(8192)
        ms = (C.c_float * 64)()
        calls = (C.c_int32 * 64)()
        n = C.c_int32()
        self.check(self.L.rb_ctx_profile_read(self.h, names, 8192, ms, calls, 64, C.byref(n)))
        ns = names.value.decode().split("\n") if n.value else []
        return {ns[i]: (float(ms[i]), int(calls[i])) for i in range(n.value)}

    def timer_start(self):
        self.check(self.L.rb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.check(self.L.rb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def host_alloc(self, nbytes, dtype=np.uint8):
        """Pinned host memory as a numpy array (kept alive by the returned object)."""
        p = C.c_void_p()
        self.check(self.L.rb_host_alloc(C.byref(p), nbytes))
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype)
        self._pinned = getattr(self, "_pinned", []) + [p]
        return arr

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.L.rb_dev_alloc(self.h, C.byref(p), nbytes))
        return p.value

    def dev_free(self, p):
        self.check(self.L.rb_dev_free(self.h, p))

    def h2d(self, dst, arr):
        arr = np.ascontiguousarray(arr)
        self.check(self.L.rb_memcpy_h2d(self.h, dst, _ptr(arr), arr.nbytes))

    def d2h(self, arr, src):
        self.check(self.L.rb_memcpy_d2h(self.h, _ptr(arr), src, arr.nbytes))

    def synth_reads_dev(self, seed, genome_len, first_read, n_reads, L, err_ppm, stride_bases, packed_dev):
        self.check(self.L.rb_synth_reads_dev(self.h, seed, genome_len, first_read, n_reads, L, err_ppm, stride_bases, packed_dev))

    def index(self, hashes, size):
        """getIndex for an array of hash values (BloomFilter.java:108-111)."""
        a = np.ascontiguousarray(hashes, dtype=np.int64)
        out = np.zeros(a.size, dtype=np.int64)
        self.check(self.L.rb_index_hashes(self.h, _ptr(a), a.size, size, _ptr(out)))
        return out

    def kmerize(self, reads, k, mode):
        """NTHashIterator family over every read: returns (fhash, rhash, base) per k-mer position."""
        n = reads.n_positions(k)
        f, r, b = (np.zeros(n, dtype=np.int64) for _ in range(3))
        self.check(self.L.rb_kmerize(self.h, *reads.args(), k, mode, _ptr(f), _ptr(r), _ptr(b)))
        return f, r, b

    def kmerize_ascii(self, seqs, k, mode):
        """The same over ASCII sequences, bit-exact for every byte value (IUPAC codes contribute their c & 0x07 row on the reverse strand)."""
        bases, off, n_reads = _ascii_chunk(seqs)
        n = int(np.maximum(np.diff(off) - k + 1, 0).sum())
        f, r, b = (np.zeros(n, dtype=np.int64) for _ in range(3))
        self.check(self.L.rb_kmerize_ascii(self.h, _ptr(bases), _ptr(off), n_reads, k, mode, _ptr(f), _ptr(r), _ptr(b)))
        return f, r, b

    def minimizers(self, reads, k, w, mode):
        """MinimizerHashIterator.next() for every window of w consecutive k-mers of every read (bloom/hash/MinimizerHashIterator.java:42-101)."""
        out = np.zeros(reads.n_positions(k + w - 1), dtype=np.int64)
        self.check(self.L.rb_minimizers(self.h, *reads.args(), k, w, mode, _ptr(out)))
        return out

    def kmerize_pairs(self, reads, k, d, mode):
        p = np.zeros(reads.n_positions(k + d), dtype=np.int64)
        self.check(self.L.rb_kmerize_pairs(self.h, *reads.args(), k, d, mode, _ptr(p)))
        return p


def encode_2bit_records(seqs):
    """ASCII sequences -> the reference's .2bit record stream (util/SeqBitsUtils.java:218-247 + 4-byte big-endian lengths)."""
    bases, off, n = _ascii_chunk(seqs)
    L = B.lib()
    size = L.rb_2bit_encode_records(_ptr(bases), _ptr(off), n, None)
    out = np.zeros(size, dtype=np.uint8)
    L.rb_2bit_encode_records(_ptr(bases), _ptr(off), n, _ptr(out))
    return out


def _ascii_chunk(seqs):
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8), off, len(bs)


class PackedReads:
    """Reads in the ingest layout of include/rnabloom_gpu.h (2-bit codes + optional unusable-base mask)."""

    def __init__(self, packed, mask, read_off, read_len, n_reads, uniform_len=0, uniform_stride=0):
        self.packed, self.mask, self.read_off, self.read_len = packed, mask, read_off, read_len
        self.n_reads, self.uniform_len, self.uniform_stride = n_reads, uniform_len, uniform_stride

    def args(self):
        return (_ptr(self.packed), _ptr(self.mask), _ptr(self.read_off), _ptr(self.read_len), self.n_reads, self.uniform_len,
                self.uniform_stride)

    def lengths(self):
        return self.read_len if self.read_len is not None else np.full(self.n_reads, self.uniform_len, dtype=np.int32)

    def n_positions(self, span):
        return int(np.maximum(self.lengths().astype(np.int64) - span + 1, 0).sum())

    def offsets(self, span):
        n = np.maximum(self.lengths().astype(np.int64) - span + 1, 0)
        return np.concatenate([[0], np.cumsum(n)])


def pack_reads(seqs, quals=None, min_qual=3, use_mask=True):
    """ASCII reads (list of str/bytes) -> PackedReads via rb_pack_reads_host."""
    L = B.lib()
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in bs])
    bases = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8)
    q = None
    if quals is not None:
        qs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in quals]
        q = np.frombuffer(b"".join(qs) + b"\0", dtype=np.uint8)
    words = int(sum((len(b) + 31) // 32 for b in bs)) + 1
    packed = np.zeros(words, dtype=np.uint64)
    mask = np.zeros(words, dtype=np.uint32)
    read_off = np.zeros(len(bs), dtype=np.int64)
    read_len = np.zeros(len(bs), dtype=np.int32)
    L.rb_pack_reads_host(_ptr(bases), _ptr(q), _ptr(off), len(bs), min_qual, _ptr(packed), _ptr(mask), _ptr(read_off), _ptr(read_len))
    return PackedReads(packed, mask if use_mask else None, read_off, read_len, len(bs))


def pack_uniform(codes, stride=None):
    """(n, L) array of 2-bit codes -> uniform-layout PackedReads (stride rounded up to a multiple of 32 bases)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    n, L = codes.shape
    stride = stride or ((L + 31) // 32) * 32
    padded = np.zeros((n, stride), dtype=np.uint64)
    padded[:, :L] = codes
    sh = (2 * (np.arange(stride) % 32)).astype(np.uint64)
    words = np.bitwise_or.reduce((padded << sh).reshape(n, stride // 32, 32), axis=2)
    return PackedReads(np.ascontiguousarray(words.reshape(-1)), None, None, None, n, L, stride)


class _Filter:
    def __init__(self, ctx, handle, owned=True):
        self.ctx, self.h, self.owned = ctx, handle, owned

    @property
    def size(self):
        return self.ctx.L.rb_filter_size(self.h)

    @property
    def num_bytes(self):
        return self.ctx.L.rb_filter_num_bytes(self.h)

    def getNumHash(self):
        return self.ctx.L.rb_filter_num_hash(self.h)

    def getPopCount(self):
        v = C.c_int64()
        self.ctx.check(self.ctx.L.rb_filter_popcount(self.h, C.byref(v)))
        return v.value

    def getFPR(self):
        v = C.c_float()
        self.ctx.check(self.ctx.L.rb_filter_fpr(self.h, C.byref(v)))
        return v.value

    def empty(self):
        self.ctx.check(self.ctx.L.rb_filter_empty(self.h))

    def destroy(self):
        if self.h and self.owned:
            self.ctx.check(self.ctx.L.rb_filter_destroy(self.h))
        self.h = None

    def download(self):
        out = np.zeros(self.num_bytes, dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_filter_download(self.h, _ptr(out), out.nbytes))
        return out

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_filter_upload(self.h, _ptr(arr), arr.nbytes))

    def save(self, desc_path, bits_path):
        self.ctx.check(self.ctx.L.rb_filter_save(self.h, str(desc_path).encode(), str(bits_path).encode()))

    def equivalent(self, other):  # BloomFilter.java:253-257
        return self.size == other.size and self.getNumHash() == other.getNumHash() and bool((self.download() == other.download()).all())

    def device_ptr(self):
        p = C.c_void_p()
        self.ctx.check(self.ctx.L.rb_filter_device_ptr(self.h, C.byref(p)))
        return p.value


def _hashes(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class BloomFilter(_Filter):
    """bloom/BloomFilter.java"""

    def __init__(self, ctx, size, numHash, k, _handle=None):
        if _handle is None:
            h = C.c_void_p()
            ctx.check(ctx.L.rb_filter_create(ctx.h, B.RB_BLOOM, size, numHash, k, C.byref(h)))
            super().__init__(ctx, h)
        else:
            super().__init__(ctx, _handle, owned=False)

    @classmethod
    def load(cls, ctx, desc_path, bits_path, k, loadBits=True):
        h = C.c_void_p()
        ctx.check(ctx.L.rb_filter_load(ctx.h, B.RB_BLOOM, str(desc_path).encode(), str(bits_path).encode(), k, int(loadBits), C.byref(h)))
        f = cls(ctx, 0, 0, k, _handle=h)
        f.owned = True
        return f

    def add(self, hashVals):
        a = _hashes(hashVals)
        self.ctx.check(self.ctx.L.rb_filter_add_hashes(self.h, _ptr(a), a.size))

    def lookup(self, hashVals):
        a = _hashes(hashVals)
        out = np.zeros(a.size, dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_filter_lookup_hashes(self.h, _ptr(a), a.size, _ptr(out)))
        return out.astype(bool)

    def lookupThenAdd(self, hashVals):
        a = _hashes(hashVals)
        out = np.zeros(a.size, dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_filter_lookup_then_add_hashes(self.h, _ptr(a), a.size, _ptr(out)))
        return out.astype(bool)

    # ---- whole sequences against this filter: the screening filter of the assembly stages (util/GraphUtils.java:627-650) ----
    def _seq_op(self, reads, mode, op):
        out = np.zeros(reads.n_reads, dtype=np.uint8) if op else None
        self.ctx.check(self.ctx.L.rb_filter_seq_op(self.h, *reads.args(), mode, op, _ptr(out)))
        return None if out is None else out.astype(bool)

    def addAllKmers(self, reads, mode=2):
        """for (Kmer kmer : kmers) bf.add(kmer.getHash()) for every read (RNABloom.java:1680); mode 0 fwd, 1 rc, 2 canonical"""
        self._seq_op(reads, mode, 0)

    def containsAllKmers(self, reads, mode=2):
        return self._seq_op(reads, mode, 1)

    def lookupAndAddAllKmers(self, reads, mode=2):
        return self._seq_op(reads, mode, 2)

    @staticmethod
    def getExpectedSize(expNumElements, fpr, numHash):
        return B.lib().rb_expected_size(expNumElements, fpr, numHash)


class KmerHistogram:
    """k-mer multiplicity histogram by hash sampling on the GPU (rb_card_*): what RNA-Bloom obtains from the external `ntcard` binary
    (RNABloom.java:5745-5768); the accessors are util/NTCardHistogram.java:28-100."""
    MAX_MULTIPLICITY = 65535

    def __init__(self, ctx, k, stranded, sample_bits=11, table_slots=1 << 26):
        self.ctx, self.k, self.sample_bits = ctx, k, sample_bits
        h = C.c_void_p()
        ctx.check(ctx.L.rb_card_create(ctx.h, k, int(stranded), sample_bits, table_slots, C.byref(h)))
        self.h = h
        self.numKmers = self.numUniqueKmers = self.numUniqueOverrepresentedKmers = 0
        self.counts = np.zeros(self.MAX_MULTIPLICITY, dtype=np.int64)

    def destroy(self):
        if self.h:
            self.ctx.check(self.ctx.L.rb_card_destroy(self.h))
            self.h = None

    def addReads(self, reads):
        n = C.c_int64()
        self.ctx.check(self.ctx.L.rb_card_add_reads(self.h, *reads.args(), C.byref(n)))
        return n.value

    def addReadsDev(self, packed_dev, n_reads, uniform_len, uniform_stride):
        n = C.c_int64()
        self.ctx.check(self.ctx.L.rb_card_add_reads_dev(self.h, packed_dev, None, None, None, n_reads, uniform_len, uniform_stride, C.byref(n)))
        return n.value

    def finish(self):
        """Reads the table: the sample's exact histogram, scaled by 2^sample_bits.  Returns (totals, raw histogram of the sample)."""
        totals = np.zeros(4, dtype=np.int64)
        raw = np.zeros(self.MAX_MULTIPLICITY + 1, dtype=np.int64)
        self.ctx.check(self.ctx.L.rb_card_histogram(self.h, _ptr(totals), _ptr(raw), self.MAX_MULTIPLICITY))
        scale = int(totals[3])
        self.numKmers = int(totals[0])                                   # F1
        self.numUniqueKmers = int(totals[2]) * scale                     # F0
        self.counts = raw[:self.MAX_MULTIPLICITY] * scale
        self.numUniqueOverrepresentedKmers = self.numUniqueKmers - int(self.counts.sum())
        return totals, raw

    def getNumSingletons(self):
        return int(self.counts[0])

    def getMinCovThreshold(self, multiplier):                            # NTCardHistogram.java:70-78
        for i in range(1, self.MAX_MULTIPLICITY):
            if multiplier * self.counts[i] > self.counts[i - 1]:
                return i
        return 0

    def getMaxCovThreshold(self, fraction):                              # :80-98
        num = int(round(fraction * self.numUniqueKmers))
        total = self.numUniqueOverrepresentedKmers
        if total >= num:
            return self.MAX_MULTIPLICITY + 1
        for i in range(self.MAX_MULTIPLICITY - 1, -1, -1):
            total += int(self.counts[i])
            if total >= num:
                return i + 1
        return self.MAX_MULTIPLICITY + 1

    def write(self, path):
        """ntcard's histogram file (F1, F0, then multiplicity <TAB> count), the input of NTCardHistogram(path) (:33-63)."""
        with open(path, "w") as fh:
            fh.write("F1\t%d\nF0\t%d\n" % (self.numKmers, self.numUniqueKmers))
            for i in range(self.MAX_MULTIPLICITY):
                fh.write("%d\t%d\n" % (i + 1, int(self.counts[i])))


class CascadingBloomFilter:
    """bloom/CascadingBloomFilter.java: numLevels Bloom filters of size / numLevels bits; add walks the levels with lookupThenAdd."""

    def __init__(self, ctx, size, numHash, k, numLevels):
        self.ctx, self.numLevels = ctx, numLevels
        h = C.c_void_p()
        ctx.check(ctx.L.rb_cascade_create(ctx.h, size, numHash, k, numLevels, C.byref(h)))
        self.h = h

    def destroy(self):
        if self.h:
            self.ctx.check(self.ctx.L.rb_cascade_destroy(self.h))
            self.h = None

    def getNumLevels(self):
        return self.numLevels

    def getBloomFilter(self, level):
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.rb_cascade_level(self.h, level, C.byref(h)))
        return BloomFilter(self.ctx, 0, 0, 0, _handle=h)

    def add(self, hashVals):
        a = _hashes(hashVals)
        self.ctx.check(self.ctx.L.rb_cascade_add_hashes(self.h, _ptr(a), len(a)))

    def lookup(self, hashVals):
        a = _hashes(hashVals)
        out = np.zeros(len(a), dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_cascade_lookup_hashes(self.h, _ptr(a), len(a), _ptr(out)))
        return out.astype(bool)

    def lookupThenAdd(self, hashVals):
        a = _hashes(hashVals)
        out = np.zeros(len(a), dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_cascade_lookup_then_add_hashes(self.h, _ptr(a), len(a), _ptr(out)))
        return out.astype(bool)

    def getFPR(self, level=None):
        return self.getBloomFilter(self.numLevels - 1 if level is None else level).getFPR()


class CountingBloomFilter(_Filter):
    """bloom/CountingBloomFilter.java"""

    def __init__(self, ctx, size, numHash, k, _handle=None):
        if _handle is None:
            h = C.c_void_p()
            ctx.check(ctx.L.rb_filter_create(ctx.h, B.RB_COUNTING, size, numHash, k, C.byref(h)))
            super().__init__(ctx, h)
        else:
            super().__init__(ctx, _handle, owned=False)

    def increment(self, hashVals):
        a = _hashes(hashVals)
        self.ctx.check(self.ctx.L.rb_cbf_increment_hashes(self.h, _ptr(a), a.size))

    def incrementAndGet(self, hashVals):
        a = _hashes(hashVals)
        out = np.zeros(a.size, dtype=np.float32)
        self.ctx.check(self.ctx.L.rb_cbf_increment_and_get_hashes(self.h, _ptr(a), a.size, _ptr(out)))
        return out

    def getCount(self, hashVals):
        a = _hashes(hashVals)
        out = np.zeros(a.size, dtype=np.float32)
        self.ctx.check(self.ctx.L.rb_cbf_count_hashes(self.h, _ptr(a), a.size, _ptr(out)))
        return out

    getExpectedSize = BloomFilter.getExpectedSize


class BloomFilterDeBruijnGraph:
    """graph/BloomFilterDeBruijnGraph.java"""

    def __init__(self, ctx, dbgbfNumBits, cbfNumBytes, pkbfNumBits, dbgbfNumHash, cbfNumHash, pkbfNumHash, k, stranded,
                 useReadPairedKmers, _handle=None):
        self.ctx = ctx
        self.k = k
        self.stranded = bool(stranded)
        if _handle is None:
            h = C.c_void_p()
            ctx.check(ctx.L.rb_graph_create(ctx.h, dbgbfNumBits, cbfNumBytes, pkbfNumBits, dbgbfNumHash, cbfNumHash, pkbfNumHash, k,
                                            int(stranded), int(useReadPairedKmers), C.byref(h)))
            self.h = h
        else:
            self.h = _handle

    @classmethod
    def load(cls, ctx, graphFile, loadDbgBits=True, loadFpkbf=True):
        h = C.c_void_p()
        ctx.check(ctx.L.rb_graph_load(ctx.h, str(graphFile).encode(), int(loadDbgBits), int(loadFpkbf), C.byref(h)))
        g = cls(ctx, 0, 0, 0, 0, 0, 0, 0, False, False, _handle=h)
        with open(graphFile) as fh:
            for line in fh:
                key, _, val = line.strip().partition(":")
                if key == "k":
                    g.k = int(val)
                elif key == "stranded":
                    g.stranded = val == "true"
        return g

    ENGINE_DIRECT, ENGINE_SLICED, ENGINE_AUTO = 0, 2, 3

    def setEngine(self, engine):
        """Execution engine of the read-level calls (include/rnabloom_gpu.h RB_ENGINE_*): same results, different HBM schedule."""
        self.ctx.check(self.ctx.L.rb_graph_set_engine(self.h, int(engine)))

    def destroy(self):
        if self.h:
            self.ctx.check(self.ctx.L.rb_graph_destroy(self.h))
            self.h = None

    def _filter(self, which, cls):
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.rb_graph_filter(self.h, which, C.byref(h)))
        return cls(self.ctx, 0, 0, self.k, _handle=h) if h.value else None

    def getDbgbf(self):
        return self._filter(B.RB_DBGBF, BloomFilter)

    def getCbf(self):
        return self._filter(B.RB_CBF, CountingBloomFilter)

    def getRpkbf(self):
        return self._filter(B.RB_RPKBF, BloomFilter)

    def getFpkbf(self):
        return self._filter(B.RB_FPKBF, BloomFilter)

    def initializePairKmersBloomFilter(self, pkbfNumBits, pkbfNumHash):
        self.ctx.check(self.ctx.L.rb_graph_init_fpkbf(self.h, pkbfNumBits, pkbfNumHash))

    def setPairedKmerDistances(self, readPairedKmersDistance, fragmentPairedKmersDistance=-1):
        self.ctx.check(self.ctx.L.rb_graph_set_distances(self.h, readPairedKmersDistance, fragmentPairedKmersDistance))
        self._d_read, self._d_frag = readPairedKmersDistance, fragmentPairedKmersDistance

    def _pair_distance(self, which):
        d = getattr(self, "_d_read" if which == B.RB_RPKBF else "_d_frag", -1)
        if d < 1:
            raise ValueError("set the paired k-mer distance first (setPairedKmerDistances)")
        return d

    def syncToHost(self, dbgbf=None, cbf=None, rpkbf=None, fpkbf=None):
        """Barrier + refresh of host mirrors (numpy uint8 arrays of the filters' byte lengths; None skips a filter): the host side of
        java/rnabloom/gpu/GpuBloomFilterDeBruijnGraph.syncToHost()."""
        self.ctx.check(self.ctx.L.rb_graph_sync_to_host(self.h, _ptr(dbgbf), _ptr(cbf), _ptr(rpkbf), _ptr(fpkbf)))

    def sync(self):
        self.ctx.check(self.ctx.L.rb_graph_sync(self.h))

    def clear(self):
        self.ctx.check(self.ctx.L.rb_graph_clear(self.h))

    # ---- bulk insert: the bodies of the *ToGraphWorker classes (RNABloom.java:364-732,1463-1539) ----
    def addReads(self, reads, flags=0):
        n = C.c_int64()
        self.ctx.check(self.ctx.L.rb_graph_add_reads(self.h, *reads.args(), flags, C.byref(n)))
        return n.value

    def addReadsAscii(self, seqs, quals=None, minQual=3, flags=0):
        bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
        off = np.zeros(len(bs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(b) for b in bs])
        bases = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8)
        q = None
        if quals is not None:
            q = np.frombuffer(b"".join(s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in quals) + b"\0", dtype=np.uint8)
        n = C.c_int64()
        self.ctx.check(self.ctx.L.rb_graph_add_reads_ascii(self.h, _ptr(bases), _ptr(q), _ptr(off), len(bs), minQual, flags, C.byref(n)))
        return n.value

    def addReads2bit(self, records, flags=0):
        """A buffer of the reference's .2bit fragment records (io/NucleotideBitsWriter.java:24-31): returns (reads, k-mers) inserted."""
        rec = np.ascontiguousarray(records, dtype=np.uint8)
        nr, nk = C.c_int64(), C.c_int64()
        self.ctx.check(self.ctx.L.rb_graph_add_reads_2bit(self.h, _ptr(rec), rec.nbytes, flags, C.byref(nr), C.byref(nk)))
        return nr.value, nk.value

    def getKmersAscii(self, seqs):
        """graph.getKmers(String) for a chunk of sequences (graph :1224-1234): (counts, fHashVals, rHashVals) of every k-mer window."""
        bases, off, n_reads = _ascii_chunk(seqs)
        n = int(np.maximum(np.diff(off) - self.k + 1, 0).sum())
        counts = np.zeros(n, dtype=np.float32)
        f, r = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        got = C.c_int64()
        self.ctx.check(self.ctx.L.rb_graph_count_reads_ascii(self.h, _ptr(bases), _ptr(off), n_reads, _ptr(counts), _ptr(f), _ptr(r), C.byref(got)))
        assert got.value == n
        return counts, f, r

    def addReadsDev(self, packed_dev, n_reads, uniform_len, uniform_stride, flags=0):
        n = C.c_int64()
        self.ctx.check(self.ctx.L.rb_graph_add_reads_dev(self.h, packed_dev, None, None, None, n_reads, uniform_len, uniform_stride, flags,
                                                         C.byref(n)))
        return n.value

    # ---- bulk lookup: graph.getKmers (graph :1224-1226) ----
    def getKmers(self, reads, want_hashes=True):
        n = reads.n_positions(self.k)
        counts = np.zeros(n, dtype=np.float32)
        fh = np.zeros(n, dtype=np.int64) if want_hashes else None
        rh = np.zeros(n, dtype=np.int64) if (want_hashes and not self.stranded) else None
        got = C.c_int64()
        self.ctx.check(self.ctx.L.rb_graph_count_reads(self.h, *reads.args(), _ptr(counts), _ptr(fh), _ptr(rh), C.byref(got)))
        assert got.value == n
        return counts, fh, rh

    def getKmersAsync(self, reads, counts, fh=None, rh=None):
        """rb_graph_count_reads_async: results land in the caller's (pinned) arrays once Context.wait(ticket) has returned."""
        got, ticket = C.c_int64(), C.c_int64()
        self.ctx.check(self.ctx.L.rb_graph_count_reads_async(self.h, *reads.args(), _ptr(counts), _ptr(fh), _ptr(rh), C.byref(got), C.byref(ticket)))
        return ticket.value

    def getKmersDev(self, packed_dev, n_reads, uniform_len, uniform_stride, counts_dev, fhash_dev=None, rhash_dev=None):
        n = C.c_int64()
        self.ctx.check(self.ctx.L.rb_graph_count_reads_dev(self.h, packed_dev, None, None, None, n_reads, uniform_len, uniform_stride,
                                                           counts_dev, fhash_dev, rhash_dev, C.byref(n)))
        return n.value

    # ---- per-hash forms ----
    def add(self, hashVals, flags=0):
        a = _hashes(hashVals)
        self.ctx.check(self.ctx.L.rb_graph_add_hashes(self.h, _ptr(a), a.size, flags))

    def addCountIfPresent(self, hashVals):
        self.add(hashVals, B.ADD_COUNT_IF_PRESENT)

    def addDbgOnly(self, hashVals):
        self.add(hashVals, B.DBG_ONLY)

    def getCount(self, hashVals):
        a = _hashes(hashVals)
        out = np.zeros(a.size, dtype=np.float32)
        self.ctx.check(self.ctx.L.rb_graph_count_hashes(self.h, _ptr(a), a.size, _ptr(out)))
        return out

    def contains(self, hashVals):
        return self.getDbgbf().lookup(hashVals)

    def getNeighborCounts(self, fHashVals, rHashVals, firstBases, lastBases, with_hashes=True):
        """Batched Kmer.getSuccessors / getPredecessors (graph/Kmer.java:213-253, CanonicalKmer.java:232-271): for every k-mer the
        graph.getCount of its 4 candidate successors and 4 candidate predecessors (A, C, G, T) -> counts[n, 2, 4] (and the candidates'
        forward / reverse hashes).  Bases are 2-bit codes A0 C1 G2 T3 of the k-mer's first and last base."""
        f = _hashes(fHashVals)
        r = None if rHashVals is None else _hashes(rHashVals)
        b0 = np.ascontiguousarray(firstBases, dtype=np.uint8)
        b1 = np.ascontiguousarray(lastBases, dtype=np.uint8)
        n = f.size
        counts = np.zeros((n, 2, 4), dtype=np.float32)
        nf = np.zeros((n, 2, 4), dtype=np.int64) if with_hashes else None
        nr = np.zeros((n, 2, 4), dtype=np.int64) if with_hashes and not self.stranded else None
        self.ctx.check(self.ctx.L.rb_graph_neighbor_counts(self.h, _ptr(f), _ptr(r), _ptr(b0), _ptr(b1), n, _ptr(counts), _ptr(nf), _ptr(nr)))
        return counts, nf, nr

    def getVariantCounts(self, fHashVals, rHashVals, firstBases, lastBases, with_hashes=True):
        """Batched Kmer.getLeftVariants / getRightVariants (graph/Kmer.java:357-405): counts[n, 2, 4] of the k-mers with base A, C, G, T in
        the first (index 0) / last (index 1) position; the entry of the k-mer's own base is the k-mer itself."""
        f = _hashes(fHashVals)
        r = None if rHashVals is None else _hashes(rHashVals)
        b0 = np.ascontiguousarray(firstBases, dtype=np.uint8)
        b1 = np.ascontiguousarray(lastBases, dtype=np.uint8)
        n = f.size
        counts = np.zeros((n, 2, 4), dtype=np.float32)
        vf = np.zeros((n, 2, 4), dtype=np.int64) if with_hashes else None
        vr = np.zeros((n, 2, 4), dtype=np.int64) if with_hashes and not self.stranded else None
        self.ctx.check(self.ctx.L.rb_graph_variant_counts(self.h, _ptr(f), _ptr(r), _ptr(b0), _ptr(b1), n, _ptr(counts), _ptr(vf), _ptr(vr)))
        return counts, vf, vr

    def getMaxCovNeighbors(self, fHashVals, rHashVals, firstBases, lastBases, minKmerCov=1.0):
        """Batched Kmer.getMaxCovSuccessor / getMaxCovPredecessor (graph/Kmer.java:301-355): (best base code or -1, its count), each [n, 2]."""
        f = _hashes(fHashVals)
        r = None if rHashVals is None else _hashes(rHashVals)
        b0 = np.ascontiguousarray(firstBases, dtype=np.uint8)
        b1 = np.ascontiguousarray(lastBases, dtype=np.uint8)
        n = f.size
        best = np.zeros((n, 2), dtype=np.int8)
        cnt = np.zeros((n, 2), dtype=np.float32)
        self.ctx.check(self.ctx.L.rb_graph_max_cov_neighbors(self.h, _ptr(f), _ptr(r), _ptr(b0), _ptr(b1), n, minKmerCov, _ptr(best), _ptr(cnt), None, None))
        return best, cnt

    def greedyExtend(self, kmers, right=True, bound=100, minKmerCov=1.0):
        """Batched GraphUtils.greedyExtendRight / greedyExtendLeft with lookahead <= 1 (util/GraphUtils.java:1961-1976): list of ASCII k-mers ->
        list of extension strings (to the left: in genome order, i.e. the prepended bases reversed)."""
        k = self.k
        code = np.full(256, 0, dtype=np.uint64)
        for ch, v in zip("ACGTUacgtu", (0, 1, 2, 3, 3, 0, 1, 2, 3, 3)):
            code[ord(ch)] = v
        n = len(kmers)
        arr = code[np.frombuffer("".join(kmers).encode(), dtype=np.uint8).reshape(n, k)]
        bits = np.zeros((n, 2), dtype=np.uint64)
        for i in range(k):
            bits[:, i >> 5] |= arr[:, i] << np.uint64(2 * (i & 31))
        f, r, _ = self.ctx.kmerize_ascii(kmers, k, B.MODE_CANON)
        ext_len = np.zeros(n, dtype=np.int32)
        ext = np.zeros((n, bound), dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_graph_greedy_extend(self.h, _ptr(np.ascontiguousarray(bits)), _ptr(f), _ptr(r), n, int(right), bound, minKmerCov,
                                                         _ptr(ext_len), _ptr(ext), None))
        nt = np.frombuffer(b"ACGT", dtype=np.uint8)
        out = []
        for i in range(n):
            s_ = bytes(nt[ext[i, :ext_len[i]]]).decode()
            out.append(s_ if right else s_[::-1])
        return out

    def addReadSingleKmerPair(self, pairHashVals):
        a = _hashes(pairHashVals)
        self.ctx.check(self.ctx.L.rb_graph_add_pair_hashes(self.h, B.RB_RPKBF, _ptr(a), a.size))

    def addFragmentSingleKmerPair(self, pairHashVals):
        a = _hashes(pairHashVals)
        self.ctx.check(self.ctx.L.rb_graph_add_pair_hashes(self.h, B.RB_FPKBF, _ptr(a), a.size))

    def lookupReadKmerPair(self, pairHashVals):
        a = _hashes(pairHashVals)
        out = np.zeros(a.size, dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_graph_lookup_pair_hashes(self.h, B.RB_RPKBF, _ptr(a), a.size, _ptr(out)))
        return out.astype(bool)

    def lookupKmerPairsOfReads(self, reads, which=B.RB_RPKBF, flags=0):
        """lookupReadKmerPair / lookupFragmentKmerPair (graph :526-532) at every pair position of every read: the test inside
        breakWith{Read,Frag}PairedKmers (util/GraphUtils.java:4184-4310).  Returns one bool per pair position."""
        d = self._pair_distance(which)
        out = np.zeros(reads.n_positions(self.k + d), dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_graph_lookup_pairs_reads(self.h, which, *reads.args(), flags, _ptr(out)))
        return out.astype(bool)

    def breakWithPairedKmers(self, reads, numPairsRequired=3, which=B.RB_RPKBF, flags=0):
        """GraphUtils.breakWithReadPairedKmers / breakWithFragPairedKmers (util/GraphUtils.java:4184-4310) for every read: the ranges [start, end) of
        k-mer indices supported by paired k-mers.  The pair look-ups of all reads come from one batched call; the segment walk is the
        reference's, per read, on the host."""
        d = self._pair_distance(which)
        found = self.lookupKmerPairsOfReads(reads, which, flags)
        off = reads.offsets(self.k + d)
        out = []
        for r in range(reads.n_reads):
            f = found[off[r]:off[r + 1]]          # f[i] = lookupKmerPair(kmers[i], kmers[i + d]), i = 0 .. lastIndex
            segments, start, end, run = [], -1, -1, 0
            for i, hit in enumerate(f):
                if hit:
                    run += 1
                    if run >= numPairsRequired:
                        if start < 0:
                            start = i - numPairsRequired + 1
                        end = i + d
                else:
                    if start >= 0 and i >= end:   # interlockDistance = 0
                        segments.append((start, end + 1))
                        start = end = -1
                    run = 0
            if start >= 0:
                segments.append((start, end + 1))
            out.append(segments)
        return out

    def lookupFragmentKmerPair(self, pairHashVals):
        a = _hashes(pairHashVals)
        out = np.zeros(a.size, dtype=np.uint8)
        self.ctx.check(self.ctx.L.rb_graph_lookup_pair_hashes(self.h, B.RB_FPKBF, _ptr(a), a.size, _ptr(out)))
        return out.astype(bool)

    def getDbgbfFPR(self):
        return self.getDbgbf().getFPR()

    def getCbfFPR(self):
        return self.getCbf().getFPR()

    def getFPR(self):  # graph :588-590
        return self.getDbgbfFPR() * self.getCbfFPR()

    def save(self, graphFile):
        self.ctx.check(self.ctx.L.rb_graph_save(self.h, str(graphFile).encode()))
