"""Kernel LOGIC check without a GPU: the unchanged kernel sources of rna-bloom_b200/csrc are compiled for the host with
tests/emu/cuda_emu.h (one fiber per CUDA thread, barriers released by a scheduler, CTAs one after the other) and the same parity tests the GPU suite runs
are pointed at that build, against the same oracle.  This is test infrastructure: the emulated library lives under tests/emu/,
the product binding never loads it, and nothing here says anything about the device's memory model or speed -- the `-m gpu`
suite on the B200 box stays the parity gate.  What it buys is that indexing / data-movement mistakes in the multi-kernel
engines are caught in the container instead of on the (scarce) GPU box.
"""
import os
import subprocess

import numpy as np
import pytest

import rnabloom_b200 as rb
from rnabloom_b200 import binding as B

import test_gpu_parity as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "librnabloom_emu.so")
CSRC = os.path.join(ROOT, "rna-bloom_b200", "csrc")


def build_emu():
    srcs = [os.path.join(CSRC, n) for n in os.listdir(CSRC)] + [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(ROOT, "include", "rnabloom_gpu.h")]
    if os.path.exists(EMU_SO) and all(os.path.getmtime(s) <= os.path.getmtime(EMU_SO) for s in srcs):
        return
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-g", "-DRB_EMU", "-x", "c++", "-I", EMU_DIR, "-fPIC", "-shared", "-pthread",
                           "-o", EMU_SO, os.path.join(CSRC, "rnabloom_gpu.cu")])


@pytest.fixture(scope="module")
def emu_lib():
    build_emu()
    keep = B._lib
    B._lib = B.bind(EMU_SO, allow_missing=True)
    yield B._lib
    B._lib = keep


@pytest.fixture(scope="module")
def ctx(emu_lib):
    c = rb.Context(0)
    yield c
    c.close()


# RB_FULL_EMU=1 runs the whole matrix (about twice as long); the default keeps one case per distinct code path so that the CPU suite stays short
FULL = os.environ.get("RB_FULL_EMU") == "1"


SLICE_ENV = {"RB_SLICE_BITS_LOG2": "17", "RB_SLICE_BYTES_LOG2": "15", "RB_SLICED_CELLS": "1", "RB_SLICED_STAGE": "1", "RB_SLICED_SUBRANGE_LOG2": "6",
             "RB_SLICE_PAIR_LOG2": "13"}
# sliced-small: tiny slices / sub-ranges so that the small test filters span hundreds of regions (every test);
# sliced-default: the production geometry; direct: validates the emulation itself (that engine is verified on the GPU)
# sliced-small-spill: RB_SLICED_SPILL=1 with region capacities far below the expected load, so that a large part of the keys takes
# the heavy-hitter path (spill list -> global table -> merged by ks_dedup / appended afterwards); off by default in the product
SPILL_ENV = {"RB_SLICED_SPILL": "1", "RB_SLICED_SUBCAP": "40", "RB_SLICED_KEYCAP": "1500"}
ONLY = {
    "sliced-small-spill": ("test_duplicates_inside_one_batch_are_linearised", "test_skewed_batch_is_redone_by_the_direct_engine"),
    "direct": ("test_getkmers_with_invalid_nucleotides", "test_screening_filter_over_whole_sequences", "test_kmer_histogram_by_hash_sampling",
               "test_minimizers_match_the_rolling_window", "test_pair_lookups_at_every_position", "test_minimizer_based_subsampling_matches_the_sequential_loop", "test_equal_length_ascii_records_and_async_counts", "test_neighbor_counts_match_oracle", "test_kmerize_ascii_is_exact_for_every_character",
               "test_cascading_bloom_filter_matches_oracle", "test_loaded_cbf_envelope_at_scale", "test_2bit_fragment_records_round_trip",
               "test_variants_max_cov_and_greedy_extension_match_oracle", "test_stage1_driver_writes_the_reference_files"),
    "sliced-default": ("test_getkmers_with_invalid_nucleotides", "test_equal_length_ascii_records_and_async_counts", "test_insert_policies_and_pair_filters", "test_kernels_are_race_free_under_tsan",
                       "test_random_geometry_matches_oracle", "test_random_uniform_layout_matches_oracle", "test_paired_slices_match_oracle",
                       "test_config3_settings_match_oracle", "test_config4_long_reads_match_oracle", "test_loaded_cbf_envelope_at_scale"),
}
SKIP = {"sliced-small": ("test_kernels_are_race_free_under_tsan", "test_neighbor_counts_match_oracle", "test_random_geometry_matches_oracle",
                         "test_random_uniform_layout_matches_oracle", "test_config3_settings_match_oracle", "test_config4_long_reads_match_oracle"),
        "sliced-small-spill": ("test_random_geometry_matches_oracle", "test_random_uniform_layout_matches_oracle")}   # the random test sets its own geometry: it runs once, under "sliced-default"   # sets its own environment; runs once (under "sliced-default")


@pytest.fixture(autouse=True, params=["sliced-small", "sliced-small-spill", "sliced-default", "direct"])
def engine(request):
    name = request.node.originalname or request.node.name
    if (request.param in ONLY and name not in ONLY[request.param]) or name in SKIP.get(request.param, ()):
        pytest.skip("not in the reduced matrix of this variant")
    keys = ["RB_ENGINE"] + list(SLICE_ENV) + list(SPILL_ENV)
    old = {k: os.environ.get(k) for k in keys}
    os.environ["RB_ENGINE"] = request.param.split("-")[0]
    for k in keys[1:]:
        os.environ.pop(k, None)
    if "small" in request.param:
        os.environ.update(SLICE_ENV)
    if "spill" in request.param:
        os.environ.update(SPILL_ENV)
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def test_duplicates_inside_one_batch_are_linearised(ctx, orc, monkeypatch):
    monkeypatch.setattr(G, "DUP_FILTER_LOG2", 24)   # 2^30 filters cost minutes of memset / compare per graph in the emulation
    G.test_duplicates_inside_one_batch_are_linearised(ctx, orc)
test_getkmers_with_invalid_nucleotides = G.test_getkmers_with_invalid_nucleotides
test_kmerize_ascii_is_exact_for_every_character = G.test_kmerize_ascii_is_exact_for_every_character
test_equal_length_ascii_records_and_async_counts = G.test_equal_length_ascii_records_and_async_counts


@pytest.mark.parametrize("mode,k,w", [(2, 25, 10), (0, 17, 1)])
def test_minimizers_match_the_rolling_window(ctx, orc, mode, k, w):
    G.test_minimizers_match_the_rolling_window(ctx, orc, mode, k, w)


@pytest.mark.parametrize("stranded,hpc", [(False, False), (True, True)])
def test_minimizer_based_subsampling_matches_the_sequential_loop(ctx, orc, stranded, hpc):
    G.test_minimizer_based_subsampling_matches_the_sequential_loop(ctx, orc, stranded, hpc)


@pytest.mark.parametrize("stranded,k,d", [(False, 25, 30), (True, 21, 7)])
def test_pair_lookups_at_every_position(ctx, orc, stranded, k, d):
    G.test_pair_lookups_at_every_position(ctx, orc, stranded, k, d)


@pytest.mark.parametrize("stranded,k,sample_bits", [(False, 25, 0), (True, 31, 2)])
def test_kmer_histogram_by_hash_sampling(ctx, orc, stranded, k, sample_bits, tmp_path):
    G.test_kmer_histogram_by_hash_sampling(ctx, orc, stranded, k, sample_bits, tmp_path)


@pytest.mark.parametrize("mode,k,num_hash", [(2, 25, 3), (0, 31, 2)])
def test_screening_filter_over_whole_sequences(ctx, orc, mode, k, num_hash):
    G.test_screening_filter_over_whole_sequences(ctx, orc, mode, k, num_hash)
test_cascading_bloom_filter_matches_oracle = G.test_cascading_bloom_filter_matches_oracle
test_2bit_fragment_records_round_trip = G.test_2bit_fragment_records_round_trip
test_stage1_driver_writes_the_reference_files = G.test_stage1_driver_writes_the_reference_files


@pytest.mark.parametrize("stranded,k", [(False, 25), (True, 31), (False, 64)])
def test_variants_max_cov_and_greedy_extension_match_oracle(ctx, orc, stranded, k):
    G.test_variants_max_cov_and_greedy_extension_match_oracle(ctx, orc, stranded, k)
test_insert_policies_and_pair_filters = G.test_insert_policies_and_pair_filters
test_neighbor_counts_match_oracle = G.test_neighbor_counts_match_oracle


# paired probe records (cbf_bytes a power of two dividing dbg_bits): q = 8, q = 5 (stranded), h_d > h_c, and the h_d < h_c case that must
# stay unpaired; both read layouts
PAIRED_CASES = [(False, 25, 3, 3, 1 << 27, 1 << 24, "ragged"), (False, 25, 3, 3, 1 << 27, 1 << 24, "uniform"), (True, 25, 3, 3, 5 << 24, 1 << 24, "uniform"),
                (False, 31, 3, 2, 1 << 25, 1 << 25, "ragged"), (False, 25, 2, 3, 1 << 27, 1 << 24, "uniform")]
if FULL:
    PAIRED_CASES += [(True, 25, 3, 3, 5 << 24, 1 << 24, "ragged"), (False, 31, 3, 2, 1 << 25, 1 << 25, "uniform"), (False, 25, 2, 3, 1 << 27, 1 << 24, "ragged")]


@pytest.mark.parametrize("stranded,k,hd,hc,dbg_bits,cbf_bytes,layout", PAIRED_CASES)
def test_paired_slices_match_oracle(ctx, orc, stranded, k, hd, hc, dbg_bits, cbf_bytes, layout):
    G.test_paired_slices_match_oracle(ctx, orc, stranded, k, hd, hc, dbg_bits, cbf_bytes, layout)


def test_loaded_cbf_envelope_at_scale(ctx, orc, monkeypatch):
    monkeypatch.setattr(G, "N_ENVELOPE_READS", 1500)
    G.test_loaded_cbf_envelope_at_scale(ctx, orc)


def test_config3_settings_match_oracle(ctx, orc, monkeypatch):
    monkeypatch.setattr(G, "N_CFG3_READS", 1200)
    monkeypatch.setattr(G, "CFG3_SIZES", (1 << 27, 1 << 24, 1 << 24))
    G.test_config3_settings_match_oracle(ctx, orc)


def test_config4_long_reads_match_oracle(ctx, orc, monkeypatch):
    monkeypatch.setattr(G, "N_CFG4_READS", 60)
    G.test_config4_long_reads_match_oracle(ctx, orc)


def test_skewed_batch_is_redone_by_the_direct_engine(ctx, orc, monkeypatch):
    monkeypatch.setattr(G, "SKEW_COPIES", 2500)   # enough for the 2179-key sub-ranges of the small test geometry, cheaper to emulate
    G.test_skewed_batch_is_redone_by_the_direct_engine(ctx, orc)




@pytest.mark.parametrize("stranded,k,hd,hc,n_reads", [(False, 25, 3, 3, 800), (True, 25, 3, 3, 120)] if FULL else [(True, 25, 3, 3, 300)])
def test_graph_add_collision_free_is_bit_exact(ctx, orc, stranded, k, hd, hc, n_reads):
    G.test_graph_add_collision_free_is_bit_exact(ctx, orc, stranded, k, hd, hc, n_reads)



# the uniform-layout cases that exercise distinct code paths (tile ends inside a read, read longer than a tile, walker fallback)
@pytest.mark.parametrize("L,stride,k,n_reads,stranded", [(150, 160, 25, 300, False), (150, 192, 25, 120, True), (3000, 3008, 25, 3, False),
                                                         (40, 64, 25, 200, False)])
def test_uniform_layout_graph_matches_oracle(ctx, orc, L, stride, k, n_reads, stranded):
    G.test_uniform_layout_graph_matches_oracle(ctx, orc, L, stride, k, n_reads, stranded)


def test_kernels_are_race_free_under_tsan(tmp_path):
    """The emulated kernels (sliced engine: uniform + variable layout, insert twice + lookup) under ThreadSanitizer.  A missing
    __syncthreads() does not change the results of the emulation (OS threads rarely interleave badly) but corrupts a full-size GPU
    run; TSan sees the unordered shared-memory accesses directly (checked: removing one barrier of TileSort::run is reported)."""
    exe = os.path.join(EMU_DIR, "race_check")
    srcs = [os.path.join(CSRC, n) for n in os.listdir(CSRC)] + [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(EMU_DIR, "race_check.cpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        b = subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fsanitize=thread", "-DRB_EMU", "-DRB_EMU_THREADS", "-I", EMU_DIR, "-pthread",
                            os.path.join(EMU_DIR, "race_check.cpp"), "-o", exe], capture_output=True, text=True)
        if b.returncode != 0 and "tsan" in b.stderr.lower():
            pytest.skip("this toolchain has no ThreadSanitizer runtime")
        assert b.returncode == 0, b.stderr[-3000:]
    env = dict(os.environ, **SLICE_ENV)
    env["RB_ENGINE"] = "sliced"
    p = subprocess.run([exe, "120"], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "race_check ok" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
    assert "WARNING: ThreadSanitizer" not in p.stderr, p.stderr[:6000]


@pytest.mark.parametrize("seed", range(int(os.environ.get("RB_FUZZ_SEEDS", "8"))))
def test_random_geometry_matches_oracle(ctx, orc, seed, monkeypatch):
    """Seeded random configurations of the sliced engine against the oracle: filter sizes that are no multiple of the slice size (or
    smaller than one slice), 1..3 hashes per filter, k from 9 to 70, ragged reads with unusable bases, random slice / sub-range
    geometry, spill path on or off; insert (twice: every k-mer present the second time), look-up of every k-mer, one more policy."""
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.integers(9, 71))
    hd, hc = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    stranded = bool(rng.integers(0, 2))
    dbg_bits = int(rng.integers(1 << 16, 1 << 25)) | 1
    cbf_bytes = int(rng.integers(1 << 20, 1 << 23)) | 1       # roomy: counters of distinct k-mers rarely collide, the comparison stays tight
    if seed % 2:   # sizes the engine pairs the probe records for (when h_d >= h_c): cbf_bytes = 2^c, dbg_bits = q * cbf_bytes
        cbf_bytes = 1 << int(rng.integers(20, 23))
        dbg_bits = cbf_bytes * int(rng.integers(1, 17))
        monkeypatch.setenv("RB_SLICE_PAIR_LOG2", str(int(rng.integers(8, 19))))
        monkeypatch.setenv("RB_SLICE_REGION_TARGET", str(int(rng.choice([2, 8, 64, 512]))))   # wide regions consumed in several passes
    for name, lo, hi in (("RB_SLICE_BITS_LOG2", 12, 22), ("RB_SLICE_BYTES_LOG2", 10, 20), ("RB_SLICED_CELLS", 0, 1),
                         ("RB_SLICED_SUBRANGE_LOG2", 4, 9)):
        monkeypatch.setenv(name, str(int(rng.integers(lo, hi + 1))))
    if rng.integers(0, 2):
        monkeypatch.setenv("RB_SLICED_SPILL", "1")
        monkeypatch.setenv("RB_SLICED_SUBCAP", str(int(rng.integers(8, 200))))
    monkeypatch.setenv("RB_ENGINE", "sliced")
    n_reads = int(rng.integers(60, 260))
    genome = "".join(rng.choice(list("ACGT"), size=max(4 * k, n_reads * 40)))
    seqs = []
    for _ in range(n_reads):
        L = int(rng.integers(1, 260))
        p = int(rng.integers(0, max(1, len(genome) - L)))
        s = list(genome[p:p + L])
        for j in np.nonzero(rng.random(len(s)) < 0.004)[0]:
            s[j] = "N"
        seqs.append("".join(s))
    g, og = G.make_graphs(ctx, orc, dbg_bits, cbf_bytes, 64, hd, hc, 1, k, stranded, False)
    modes = [G.MODE_FWD, G.MODE_RC] if stranded else [G.MODE_CANON]
    bases = G.all_bases(orc, [s.replace("N", "A") for s in seqs], k, modes)
    half = n_reads // 2
    for rnd, (chunk, flags, oflags) in enumerate(((seqs, 0, 0), (seqs[:half], rb.REVCOMP if stranded else 0, G.F_REVCOMP if stranded else 0),
                                                   (seqs[half:], rb.ADD_COUNT_IF_PRESENT, G.F_ADD_COUNT_IF_PRESENT), (seqs[:20], rb.DBG_ONLY, G.F_DBG_ONLY))):
        for s in chunk:
            og.add_read(s, flags=oflags)
        g.addReads(rb.pack_reads(chunk), flags=flags)
        assert (g.getDbgbf().download() == og.dbgbf()).all(), "dbgbf differs after round %d" % rnd
    if og.cbf().max() <= 15:      # strictly inside the deterministic MiniFloat range (an increment attempted at 16 is already a coin flip)
        diff = np.nonzero(g.getCbf().download() != og.cbf())[0]
        if len(diff):
            # order dependence the reference has itself: k-mers that share a counter, and k-mers that share a dbgbf bit with another
            # k-mer (whether such a k-mer is "present" the first time it is seen depends on which of the two came first)
            allowed, _ = G.counters_that_may_differ(bases, k, hc, cbf_bytes)
            bits = G.np_slots(bases, k, hd, dbg_bits)
            uniq, cnt = np.unique(bits.reshape(-1), return_counts=True)
            fp_prone = np.isin(bits, uniq[cnt > 1]).any(axis=1)
            allowed |= set(int(x) for x in G.np_slots(bases[fp_prone], k, hc, cbf_bytes).reshape(-1))
            assert set(diff.tolist()) <= allowed, "cbf differs on counters of k-mers that share neither a counter nor a dbgbf bit"
        counts, fh, rh = g.getKmers(rb.pack_reads(seqs))
        want = [og.count_seq(s) for s in seqs]
        assert (fh == np.concatenate([w[1] for w in want])).all()
        if not stranded:
            assert (rh == np.concatenate([w[2] for w in want])).all()
        if not len(diff):
            assert (counts == np.concatenate([w[0] for w in want])).all()
    g.destroy(), og.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("RB_FUZZ_SEEDS", "2"))))
def test_random_uniform_layout_matches_oracle(ctx, orc, seed, monkeypatch):
    """The uniform (fixed-length, strided) ingest with random read length, stride, k and engine geometry: the XOR-prefix k-merizer's
    tile / span arithmetic (tiles ending inside reads, reads longer than a tile, layouts it must hand to the rolling walker)."""
    rng = np.random.default_rng(5000 + seed)
    k = int(rng.integers(15, 66))   # shorter k-mers repeat by chance in the synthetic genome and push counters to 16, where MiniFloat flips coins
    L = int(rng.integers(k, k + 500))
    stride = ((L + 31) // 32 + int(rng.integers(0, 3))) * 32
    n_reads = max(2, int(rng.integers(20000, 90000)) // (L - k + 1))
    for name, lo, hi in (("RB_SLICE_BITS_LOG2", 14, 22), ("RB_SLICE_BYTES_LOG2", 12, 20), ("RB_SLICED_CELLS", 0, 1),
                         ("RB_SLICED_SUBRANGE_LOG2", 4, 9)):
        monkeypatch.setenv(name, str(int(rng.integers(lo, hi + 1))))
    monkeypatch.setenv("RB_ENGINE", "sliced")
    G.test_uniform_layout_graph_matches_oracle(ctx, orc, L, stride, k, n_reads, bool(rng.integers(0, 2)))
