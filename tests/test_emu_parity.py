"""Kernel LOGIC check without a GPU: the unchanged kernel sources of rna-bloom_b200/csrc are compiled for the host with
tests/emu/cuda_emu.h (one OS thread per CUDA thread, CTAs one after the other) and the same parity tests the GPU suite runs
are pointed at that build, against the same oracle.  This is test infrastructure: the emulated library lives under tests/emu/,
the product binding never loads it, and nothing here says anything about the device's memory model or speed -- the `-m gpu`
suite on the B200 box stays the parity gate.  What it buys is that indexing / data-movement mistakes in the multi-kernel
engines are caught in the container instead of on the (scarce) GPU box.
"""
import os
import subprocess

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


SLICE_ENV = {"RB_SLICE_BITS_LOG2": "17", "RB_SLICE_BYTES_LOG2": "15", "RB_SLICE_RAISE_LOG2": "14", "RB_SLICE_TABLE_LOG2": "10"}


@pytest.fixture(autouse=True, params=["sliced", "sliced-default-slices", "direct"])
def engine(request):
    """sliced: tiny slices so that the small test filters span hundreds of regions; sliced-default-slices: the production geometry."""
    name = request.node.originalname or request.node.name
    if request.param == "direct" and name not in ("test_getkmers_with_invalid_nucleotides", "test_insert_policies_and_pair_filters"):
        pytest.skip("the direct engine is only here to validate the emulation itself (it is verified on the GPU)")
    keys = ["RB_ENGINE"] + list(SLICE_ENV)
    old = {k: os.environ.get(k) for k in keys}
    os.environ["RB_ENGINE"] = request.param.split("-")[0]
    if request.param == "sliced":
        os.environ.update(SLICE_ENV)
    else:
        for k in SLICE_ENV:
            os.environ.pop(k, None)
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


test_duplicates_inside_one_batch_are_linearised = G.test_duplicates_inside_one_batch_are_linearised
test_getkmers_with_invalid_nucleotides = G.test_getkmers_with_invalid_nucleotides
test_subbatching_and_claim_table_recycling_do_not_change_results = G.test_subbatching_and_claim_table_recycling_do_not_change_results
test_insert_policies_and_pair_filters = G.test_insert_policies_and_pair_filters
test_loaded_filter_dbgbf_exact_cbf_within_envelope = G.test_loaded_filter_dbgbf_exact_cbf_within_envelope


@pytest.mark.parametrize("stranded,k,hd,hc,n_reads", [(False, 25, 3, 3, 800), (True, 25, 3, 3, 120), (False, 17, 3, 2, 120), (True, 64, 1, 4, 120)])
def test_graph_add_collision_free_is_bit_exact(ctx, orc, stranded, k, hd, hc, n_reads):
    G.test_graph_add_collision_free_is_bit_exact(ctx, orc, stranded, k, hd, hc, n_reads)
