"""ctypes binding of oracle/liboracle.so (rnabloom_oracle.c).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

MODE_FWD, MODE_RC, MODE_CANON = 0, 1, 2
F_REVCOMP, F_ADD_COUNT_IF_PRESENT, F_DBG_ONLY, F_STORE_READ_PAIRS, F_STORE_FRAG_PAIRS = 1, 2, 4, 8, 16


def build(force=False):
    src = os.path.join(_HERE, "rnabloom_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = C.c_void_p


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def load():
    lib = C.CDLL(build())
    vp, i64, i32, u64, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_float
    sigs = {
        "orc_init": (None, []),
        "orc_seed": (u64, [i32]),
        "orc_mstab": (u64, [i32, i32]),
        "orc_ntp64": (i64, [vp, i32, i32]),
        "orc_ntp64rc": (i64, [vp, i32, i32]),
        "orc_ntm64": (None, [i64, vp, i32, i32]),
        "orc_roll_fwd": (i64, [i64, i32, i32, i32]),
        "orc_roll_rc": (i64, [i64, i32, i32, i32]),
        "orc_combine": (i64, [i64, i64]),
        "orc_kmer_hashes": (i64, [vp, i32, i32, i32, i32, vp, vp, vp]),
        "orc_pair_hashes": (i64, [vp, i32, i32, i32, i32, i32, vp, vp, vp]),
        "orc_index": (i64, [i64, i64]),
        "orc_bf_create": (vp, [i64, i32, i32]),
        "orc_bf_destroy": (None, [vp]),
        "orc_bf_empty": (None, [vp]),
        "orc_bf_bytes": (vp, [vp]),
        "orc_bf_nbytes": (i64, [vp]),
        "orc_bf_size": (i64, [vp]),
        "orc_bf_add": (None, [vp, vp]),
        "orc_bf_lookup": (i32, [vp, vp]),
        "orc_bf_lookup_then_add": (i32, [vp, vp]),
        "orc_bf_add1": (None, [vp, i64]),
        "orc_bf_lookup1": (i32, [vp, i64]),
        "orc_bf_lookup_then_add1": (i32, [vp, i64]),
        "orc_bf_popcount": (i64, [vp]),
        "orc_bf_fpr": (f32, [vp]),
        "orc_expected_size": (i64, [i64, f32, i32]),
        "orc_seed_rng": (None, [u64]),
        "orc_minifloat_increment": (C.c_int8, [C.c_int8]),
        "orc_minifloat_to_float": (f32, [C.c_int8]),
        "orc_cbf_create": (vp, [i64, i32, i32]),
        "orc_cbf_destroy": (None, [vp]),
        "orc_cbf_empty": (None, [vp]),
        "orc_cbf_bytes": (vp, [vp]),
        "orc_cbf_size": (i64, [vp]),
        "orc_cbf_increment": (C.c_int8, [vp, vp]),
        "orc_cbf_increment_and_get": (f32, [vp, vp]),
        "orc_cbf_get_count": (f32, [vp, vp]),
        "orc_cbf_increment1": (C.c_int8, [vp, i64]),
        "orc_cbf_get_count1": (f32, [vp, i64]),
        "orc_cbf_popcount": (i64, [vp]),
        "orc_cbf_fpr": (f32, [vp]),
        "orc_cascade_create": (vp, [i64, i32, i32, i32]),
        "orc_cascade_destroy": (None, [vp]),
        "orc_cascade_level": (vp, [vp, i32]),
        "orc_cascade_add1": (None, [vp, i64]),
        "orc_cascade_lookup1": (i32, [vp, i64]),
        "orc_cascade_lookup_then_add1": (i32, [vp, i64]),
        "orc_graph_create": (vp, [i64, i64, i64, i32, i32, i32, i32, i32, i32]),
        "orc_graph_init_fpkbf": (None, [vp, i64, i32]),
        "orc_graph_destroy": (None, [vp]),
        "orc_graph_dbgbf": (vp, [vp]),
        "orc_graph_cbf": (vp, [vp]),
        "orc_graph_rpkbf": (vp, [vp]),
        "orc_graph_fpkbf": (vp, [vp]),
        "orc_graph_set_distances": (None, [vp, i32, i32]),
        "orc_graph_add": (None, [vp, vp]),
        "orc_graph_add_count_if_present": (None, [vp, vp]),
        "orc_graph_add_dbg_only": (None, [vp, vp]),
        "orc_graph_contains": (i32, [vp, vp]),
        "orc_graph_get_count": (f32, [vp, vp]),
        "orc_graph_neighbors": (None, [vp, i64, i64, i32, i32, vp, vp, vp]),
        "orc_graph_variants": (None, [vp, vp, i32, vp, vp, vp]),
        "orc_graph_greedy_extend": (i32, [vp, vp, i32, i32, f32, vp]),
        "orc_segment": (i32, [vp, vp, i32, i32, i32, vp, vp]),
        "orc_graph_add_segment": (i64, [vp, vp, i32, i32, i32]),
        "orc_graph_add_read": (i64, [vp, vp, vp, i32, i32, i32]),
        "orc_graph_count_seq": (i64, [vp, vp, i32, i32, vp, vp, vp]),
        "orc_synth_read": (None, [u64, u64, u64, i32, C.c_uint32, vp]),
        "orc_synth_reads": (None, [u64, u64, u64, u64, i32, C.c_uint32, vp]),
        "orc_graph_run_mt": (i64, [vp, vp, i64, i32, i32, i32, i32, vp]),
        "orc_graph_run_mt_ragged": (i64, [vp, vp, vp, i64, i32, i32, i32, i32, vp]),
        "orc_2bit_record": (i64, [vp, i32, vp]),
        "orc_2bit_decode": (None, [vp, i32, vp]),
        "orc_synth_long_len": (i32, [u64, u64]),
        "orc_synth_long_read": (i32, [u64, u64, u64, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
        "orc_synth_long_reads": (i64, [u64, u64, u64, u64, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.orc_init()
    return lib


def _buf(seq):
    if isinstance(seq, str):
        seq = seq.encode("latin-1")
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


class Oracle:
    """Convenience wrapper used by the tests; every method maps 1:1 onto an orc_* function."""

    def __init__(self):
        self.lib = load()

    # ---- hashing -------------------------------------------------------------------------
    def ntp64(self, seq, k, start=0):
        b = _buf(seq)
        return self.lib.orc_ntp64(b.ctypes.data, k, start)

    def ntp64rc(self, seq, k, start=0):
        b = _buf(seq)
        return self.lib.orc_ntp64rc(b.ctypes.data, k, start)

    def ntm64(self, base, k, m):
        hv = np.zeros(m, dtype=np.int64)
        self.lib.orc_ntm64(int(base), hv.ctypes.data, k, m)
        return hv

    def combine(self, a, b):
        return self.lib.orc_combine(int(a), int(b))

    def index(self, h, size):
        return self.lib.orc_index(int(h), int(size))

    def kmer_hashes(self, seq, k, mode, start=0, end=None):
        b = _buf(seq)
        end = len(b) if end is None else end
        n = max(0, end - start - k + 1)
        fh, rh, base = (np.zeros(n, dtype=np.int64) for _ in range(3))
        got = self.lib.orc_kmer_hashes(b.ctypes.data, start, end, k, mode, fh.ctypes.data, rh.ctypes.data, base.ctypes.data)
        assert got == n
        return fh, rh, base

    def pair_hashes(self, seq, k, d, mode, start=0, end=None):
        b = _buf(seq)
        end = len(b) if end is None else end
        n = max(0, end - start - k - d + 1)
        L, R, P = (np.zeros(n, dtype=np.int64) for _ in range(3))
        got = self.lib.orc_pair_hashes(b.ctypes.data, start, end, k, d, mode, L.ctypes.data, R.ctypes.data, P.ctypes.data)
        assert got == n
        return L, R, P

    def segment(self, seq, qual, k, min_qual=3):
        b = _buf(seq)
        q = None if qual is None else _buf(qual)
        cap = len(b) // max(k, 1) + 2
        st = np.zeros(cap, dtype=np.int32)
        en = np.zeros(cap, dtype=np.int32)
        n = self.lib.orc_segment(b.ctypes.data, None if q is None else q.ctypes.data, len(b), k, min_qual, st.ctypes.data, en.ctypes.data)
        return list(zip(st[:n].tolist(), en[:n].tolist()))

    def synth_reads(self, seed, genome_len, first, n, L, err_ppm):
        out = np.zeros((n, L), dtype=np.uint8)
        self.lib.orc_synth_reads(seed, genome_len, first, n, L, err_ppm, out.ctypes.data)
        return out

    def synth_long_reads(self, seed, genome_len, first, n, sub_ppm, ins_ppm, del_ppm):
        """ONT-like ragged reads (BASELINE configs[4]): (ASCII bases concatenated, offsets[n + 1])."""
        lens = np.array([self.lib.orc_synth_long_len(seed, first + r) for r in range(n)], dtype=np.int64)
        out = np.zeros(int(lens.sum()) + 8, dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.int64)
        total = self.lib.orc_synth_long_reads(seed, genome_len, first, n, sub_ppm, ins_ppm, del_ppm, out.ctypes.data, off.ctypes.data)
        return out[:total], off

    # ---- raw buffer views ----------------------------------------------------------------
    def bf_array(self, bf):
        n = self.lib.orc_bf_nbytes(bf)
        return np.ctypeslib.as_array(C.cast(self.lib.orc_bf_bytes(bf), C.POINTER(C.c_uint8)), shape=(n,))

    def cbf_array(self, cbf):
        n = self.lib.orc_cbf_size(cbf)
        return np.ctypeslib.as_array(C.cast(self.lib.orc_cbf_bytes(cbf), C.POINTER(C.c_uint8)), shape=(n,))


class OracleGraph:
    """BloomFilterDeBruijnGraph restated on the CPU (sequential semantics)."""

    def __init__(self, orc, dbg_bits, cbf_bytes, pkbf_bits, hd, hc, hp, k, stranded, use_read_pairs):
        self.o = orc
        self.lib = orc.lib
        self.k = k
        self.g = self.lib.orc_graph_create(dbg_bits, cbf_bytes, pkbf_bits, hd, hc, hp, k, int(stranded), int(use_read_pairs))

    def close(self):
        if self.g:
            self.lib.orc_graph_destroy(self.g)
            self.g = None

    def set_distances(self, d_read, d_frag):
        self.lib.orc_graph_set_distances(self.g, d_read, d_frag)

    def init_fpkbf(self, bits, hp):
        self.lib.orc_graph_init_fpkbf(self.g, bits, hp)

    def add_read(self, seq, qual=None, min_qual=3, flags=0):
        b = _buf(seq)
        q = None if qual is None else _buf(qual)
        return self.lib.orc_graph_add_read(self.g, b.ctypes.data, None if q is None else q.ctypes.data, len(b), min_qual, flags)

    def add_segment(self, seq, start, end, flags=0):
        b = _buf(seq)
        return self.lib.orc_graph_add_segment(self.g, b.ctypes.data, start, end, flags)

    def count_seq(self, seq, start=0, end=None):
        b = _buf(seq)
        end = len(b) if end is None else end
        n = max(0, end - start - self.k + 1)
        counts = np.zeros(n, dtype=np.float32)
        fh = np.zeros(n, dtype=np.int64)
        rh = np.zeros(n, dtype=np.int64)
        self.lib.orc_graph_count_seq(self.g, b.ctypes.data, start, end, counts.ctypes.data, fh.ctypes.data, rh.ctypes.data)
        return counts, fh, rh

    def neighbors(self, fh, rh, char_out, successors):
        """counts, forward and reverse hashes of the 4 candidate neighbours (A, C, G, T) of one k-mer (graph/Kmer.java:213-253)."""
        c = np.zeros(4, dtype=np.float32)
        f = np.zeros(4, dtype=np.int64)
        r = np.zeros(4, dtype=np.int64)
        self.lib.orc_graph_neighbors(self.g, int(fh), int(rh), int(char_out), int(successors), c.ctypes.data, f.ctypes.data, r.ctypes.data)
        return c, f, r

    def variants(self, kmer, side):
        """counts, forward and reverse hashes of the 4 k-mers with A, C, G, T in the first (side 0) / last (side 1) position."""
        b = _buf(kmer)
        c = np.zeros(4, dtype=np.float32)
        f = np.zeros(4, dtype=np.int64)
        r = np.zeros(4, dtype=np.int64)
        self.lib.orc_graph_variants(self.g, b.ctypes.data, int(side), c.ctypes.data, f.ctypes.data, r.ctypes.data)
        return c, f, r

    def greedy_extend(self, kmer, right=True, bound=100, min_cov=1.0):
        b = _buf(kmer)
        out = np.zeros(bound, dtype=np.uint8)
        n = self.lib.orc_graph_greedy_extend(self.g, b.ctypes.data, int(right), bound, min_cov, out.ctypes.data)
        s = bytes(out[:n]).decode()
        return s if right else s[::-1]

    def dbgbf(self):
        return self.o.bf_array(self.lib.orc_graph_dbgbf(self.g))

    def cbf(self):
        return self.o.cbf_array(self.lib.orc_graph_cbf(self.g))

    def rpkbf(self):
        return self.o.bf_array(self.lib.orc_graph_rpkbf(self.g))

    def fpkbf(self):
        return self.o.bf_array(self.lib.orc_graph_fpkbf(self.g))

    def run_mt_ragged(self, bases, off, flags, lookup, nthreads):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        cs = C.c_double(0)
        max_len = int(np.diff(off).max()) if len(off) > 1 else 0
        km = self.lib.orc_graph_run_mt_ragged(self.g, bases.ctypes.data, off.ctypes.data, len(off) - 1, max_len, flags, int(lookup), nthreads,
                                              C.addressof(cs))
        return km, cs.value

    def run_mt(self, reads, flags, lookup, nthreads):
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n, L = reads.shape
        cs = C.c_double(0)
        km = self.lib.orc_graph_run_mt(self.g, reads.ctypes.data, n, L, flags, int(lookup), nthreads, C.addressof(cs))
        return km, cs.value
