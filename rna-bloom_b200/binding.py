"""ctypes binding of librnabloom_gpu.so -- one Python function per C-ABI entry point of include/rnabloom_gpu.h."""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "librnabloom_gpu.so")
_SRC = [os.path.join(_HERE, "csrc", n) for n in ("rnabloom_gpu.cu", "rb_kernels.cuh", "rb_device.cuh", "rb_sliced.cuh",
                                                 "rb_sliced_host.inl", "rb_mgraph_host.inl")] + [
    os.path.join(_ROOT, "include", "rnabloom_gpu.h")]

RB_BLOOM, RB_COUNTING = 0, 1
RB_DBGBF, RB_CBF, RB_RPKBF, RB_FPKBF = 0, 1, 2, 3
MODE_FWD, MODE_RC, MODE_CANON = 0, 1, 2
REVCOMP, ADD_COUNT_IF_PRESENT, DBG_ONLY, STORE_READ_PAIRS, STORE_FRAG_PAIRS, PAIRS_EXISTING_ONLY = 1, 2, 4, 8, 16, 32


class RBError(RuntimeError):
    """Non-zero return of a C-ABI call (the JNI shim turns the same codes into RuntimeException)."""

    def __init__(self, code, msg):
        super().__init__("rnabloom_gpu error %d: %s" % (code, msg))
        self.code = code


def lib_path():
    return _SO


def build_library(force=False, verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRC)
    if stale:
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-o", _SO, _SRC[0]]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd, cwd=_HERE)
    return _SO


def declared_symbols():
    """Every RB_API function the header declares (used by the CPU test that checks the exports)."""
    with open(os.path.join(_ROOT, "include", "rnabloom_gpu.h")) as fh:
        text = fh.read()
    return sorted(set(re.findall(r"RB_API\s+[\w\s\*]+?\b(rb_\w+)\s*\(", text)))


_lib = None


def lib():
    """Load the library (must have been built: __graft_entry__.build() or build_library())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RBError(-3, "librnabloom_gpu.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    _lib = bind(_SO)
    return _lib


def bind(path, allow_missing=False):
    """dlopen `path` and declare the signature of every entry point.  allow_missing is for the host-emulation build of the
    kernels that tests/test_emu_parity.py uses (it has no sharded pipeline); the product library must export everything."""
    L = C.CDLL(path)
    vp, i64, i32, u32, u64, f32, cp = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64, C.c_float, C.c_char_p
    reads = [vp, vp, vp, vp, i64, i32, i64]  # packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride
    sigs = {
        "rb_version": (i32, []),
        "rb_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "rb_ctx_destroy": (i32, [vp]),
        "rb_last_error": (cp, [vp]),
        "rb_ctx_set_stream": (i32, [vp, vp]),
        "rb_ctx_sync": (i32, [vp]),
        "rb_ctx_set_rng_seed": (i32, [vp, u64]),
        "rb_ctx_set_subbatch_kmers": (i32, [vp, i64]),
        "rb_ctx_kernel_launches": (i64, [vp]),
        "rb_ctx_profile_enable": (i32, [vp, i32]),
        "rb_ctx_profile_read": (i32, [vp, vp, i64, vp, vp, i32, C.POINTER(i32)]),
        "rb_timer_start": (i32, [vp]),
        "rb_timer_stop": (i32, [vp, C.POINTER(f32)]),
        "rb_host_alloc": (i32, [C.POINTER(vp), i64]),
        "rb_host_free": (i32, [vp]),
        "rb_dev_alloc": (i32, [vp, C.POINTER(vp), i64]),
        "rb_dev_free": (i32, [vp, vp]),
        "rb_memcpy_h2d": (i32, [vp, vp, vp, i64]),
        "rb_memcpy_d2h": (i32, [vp, vp, vp, i64]),
        "rb_expected_size": (i64, [i64, f32, i32]),
        "rb_minifloat_to_float": (f32, [C.c_int8]),
        "rb_pack_reads_host": (i64, [vp, vp, vp, i64, i32, vp, vp, vp, vp]),
        "rb_kmer_offsets": (i64, [vp, i64, i32, i32, vp]),
        "rb_filter_create": (i32, [vp, i32, i64, i32, i32, C.POINTER(vp)]),
        "rb_filter_destroy": (i32, [vp]),
        "rb_filter_empty": (i32, [vp]),
        "rb_filter_size": (i64, [vp]),
        "rb_filter_num_bytes": (i64, [vp]),
        "rb_filter_num_hash": (i32, [vp]),
        "rb_filter_device_ptr": (i32, [vp, C.POINTER(vp)]),
        "rb_filter_add_hashes": (i32, [vp, vp, i64]),
        "rb_filter_lookup_hashes": (i32, [vp, vp, i64, vp]),
        "rb_filter_lookup_then_add_hashes": (i32, [vp, vp, i64, vp]),
        "rb_cbf_increment_hashes": (i32, [vp, vp, i64]),
        "rb_cbf_increment_and_get_hashes": (i32, [vp, vp, i64, vp]),
        "rb_cbf_count_hashes": (i32, [vp, vp, i64, vp]),
        "rb_filter_popcount": (i32, [vp, C.POINTER(i64)]),
        "rb_filter_fpr": (i32, [vp, C.POINTER(f32)]),
        "rb_filter_download": (i32, [vp, vp, i64]),
        "rb_filter_upload": (i32, [vp, vp, i64]),
        "rb_filter_save": (i32, [vp, cp, cp]),
        "rb_filter_load": (i32, [vp, i32, cp, cp, i32, i32, C.POINTER(vp)]),
        "rb_cascade_create": (i32, [vp, i64, i32, i32, i32, C.POINTER(vp)]),
        "rb_cascade_destroy": (i32, [vp]),
        "rb_cascade_level": (i32, [vp, i32, C.POINTER(vp)]),
        "rb_cascade_add_hashes": (i32, [vp, vp, i64]),
        "rb_cascade_lookup_hashes": (i32, [vp, vp, i64, vp]),
        "rb_cascade_lookup_then_add_hashes": (i32, [vp, vp, i64, vp]),
        "rb_index_hashes": (i32, [vp, vp, i64, i64, vp]),
        "rb_kmerize": (i32, [vp] + reads + [i32, i32, vp, vp, vp]),
        "rb_kmerize_pairs": (i32, [vp] + reads + [i32, i32, i32, vp]),
        "rb_nccl_unique_id": (i32, [vp, i64]),
        "rb_mgraph_create_nccl": (i32, [vp, i32, i32, vp, i64, i64, i32, i32, i32, i32, i64, C.POINTER(vp)]),
        "rb_mgraph_create": (i32, [vp, i32, i32, vp, i64, i64, i32, i32, i32, i32, i64, C.POINTER(vp)]),
        "rb_mgraph_destroy": (i32, [vp]),
        "rb_mgraph_layout": (i32, [vp, C.POINTER(i64)]),
        "rb_mgraph_filter": (i32, [vp, i32, C.POINTER(vp)]),
        "rb_mgraph_stats": (i32, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "rb_mgraph_peer_to_peer": (i32, [vp]),
        "rb_mgraph_add_round_dev": (i32, [vp] + reads + [u32, C.POINTER(i64)]),
        "rb_mgraph_count_round_dev": (i32, [vp] + reads + [vp, vp, vp, C.POINTER(i64)]),
        "rb_graph_create": (i32, [vp, i64, i64, i64, i32, i32, i32, i32, i32, i32, C.POINTER(vp)]),
        "rb_graph_destroy": (i32, [vp]),
        "rb_graph_init_fpkbf": (i32, [vp, i64, i32]),
        "rb_graph_set_distances": (i32, [vp, i32, i32]),
        "rb_graph_filter": (i32, [vp, i32, C.POINTER(vp)]),
        "rb_graph_clear": (i32, [vp]),
        "rb_graph_set_engine": (i32, [vp, i32]),
        "rb_graph_add_reads": (i32, [vp] + reads + [u32, C.POINTER(i64)]),
        "rb_graph_add_reads_dev": (i32, [vp] + reads + [u32, C.POINTER(i64)]),
        "rb_graph_add_reads_ascii": (i32, [vp, vp, vp, vp, i64, i32, u32, C.POINTER(i64)]),
        "rb_graph_count_reads": (i32, [vp] + reads + [vp, vp, vp, C.POINTER(i64)]),
        "rb_2bit_record_bytes": (i64, [i32]),
        "rb_2bit_encode_records": (i64, [vp, vp, i64, vp]),
        "rb_2bit_index_records": (i64, [vp, i64, vp, vp, i64]),
        "rb_graph_add_reads_2bit": (i32, [vp, vp, i64, u32, C.POINTER(i64), C.POINTER(i64)]),
        "rb_graph_count_reads_ascii": (i32, [vp, vp, vp, i64, vp, vp, vp, C.POINTER(i64)]),
        "rb_kmerize_ascii": (i32, [vp, vp, vp, i64, i32, i32, vp, vp, vp]),
        "rb_graph_count_reads_dev": (i32, [vp] + reads + [vp, vp, vp, C.POINTER(i64)]),
        "rb_graph_count_reads_async": (i32, [vp] + reads + [vp, vp, vp, C.POINTER(i64), C.POINTER(i64)]),
        "rb_ctx_wait": (i32, [vp, i64]),
        "rb_filter_seq_op": (i32, [vp] + reads + [i32, i32, vp]),
        "rb_minimizers": (i32, [vp] + reads + [i32, i32, i32, vp]),
        "rb_graph_lookup_pairs_reads": (i32, [vp, i32] + reads + [u32, vp]),
        "rb_card_create": (i32, [vp, i32, i32, i32, i64, C.POINTER(vp)]),
        "rb_card_destroy": (i32, [vp]),
        "rb_card_add_reads": (i32, [vp] + reads + [C.POINTER(i64)]),
        "rb_card_add_reads_dev": (i32, [vp] + reads + [C.POINTER(i64)]),
        "rb_card_histogram": (i32, [vp, vp, vp, i32]),
        "rb_graph_add_hashes": (i32, [vp, vp, i64, u32]),
        "rb_graph_count_hashes": (i32, [vp, vp, i64, vp]),
        "rb_graph_add_pair_hashes": (i32, [vp, i32, vp, i64]),
        "rb_graph_lookup_pair_hashes": (i32, [vp, i32, vp, i64, vp]),
        "rb_graph_neighbor_counts": (i32, [vp, vp, vp, vp, vp, i64, vp, vp, vp]),
        "rb_graph_variant_counts": (i32, [vp, vp, vp, vp, vp, i64, vp, vp, vp]),
        "rb_graph_max_cov_neighbors": (i32, [vp, vp, vp, vp, vp, i64, f32, vp, vp, vp, vp]),
        "rb_graph_greedy_extend": (i32, [vp, vp, vp, vp, i64, i32, i32, f32, vp, vp, vp]),
        "rb_graph_sync": (i32, [vp]),
        "rb_graph_sync_to_host": (i32, [vp, vp, vp, vp, vp]),
        "rb_graph_save": (i32, [vp, cp]),
        "rb_graph_load": (i32, [vp, cp, i32, i32, C.POINTER(vp)]),
        "rb_synth_reads_dev": (i32, [vp, u64, u64, u64, i64, i32, u32, i64, vp]),
        "rb_synth_long_read_len": (i32, [u64, u64]),
        "rb_synth_long_reads_dev": (i32, [vp, u64, u64, u64, i64, u32, u32, u32, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        if allow_missing and not hasattr(L, name):
            continue
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._sigs = sigs
    return L
