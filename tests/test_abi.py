"""CPU-only checks of the drop-in boundary: the library builds/loads, exports every symbol the header declares, the
pure-host helpers agree with the oracle, and the product never routes through the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rnabloom_b200 as rb
from rnabloom_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    rb.build_library()
    return rb.lib()


def test_every_declared_symbol_is_exported_and_bound(L):
    names = B.declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(L, n), "header declares %s but the library does not export it" % n
        assert n in L._sigs, "python binding lacks %s" % n


def test_no_device_means_loud_failure(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rb.RBError) as ei:
        rb.Context(0)
    assert ei.value.code == -3


def test_host_helpers_match_oracle(L, orc):
    for n, fpr, h in ((10 ** 6, 0.01, 3), (5 * 10 ** 9, 0.005, 2), (123456789, 0.05, 1)):
        assert L.rb_expected_size(n, fpr, h) == orc.lib.orc_expected_size(n, fpr, h)
    for b in range(-128, 128):
        assert L.rb_minifloat_to_float(b) == orc.lib.orc_minifloat_to_float(b)
    lens = np.array([10, 25, 150, 24, 0, 300], dtype=np.int32)
    off = np.zeros(7, dtype=np.int64)
    tot = L.rb_kmer_offsets(lens.ctypes.data, 6, 0, 25, off.ctypes.data)
    assert off.tolist() == [0, 0, 1, 127, 127, 127, 403] and tot == 403


def test_host_packing_layout(L, kat):
    e = kat["segments"][3]
    pr = rb.pack_reads([e["seq"], "ACGTN", e["seq"]], [e["qual"], "IIIII", e["qual"]], min_qual=e["min_qual"])
    assert pr.read_off.tolist() == [0, ((len(e["seq"]) + 31) // 32) * 32, ((len(e["seq"]) + 31) // 32) * 32 + 32]
    code = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}
    for i, ch in enumerate(e["seq"]):
        w, sh = pr.packed[i >> 5], 2 * (i & 31)
        m = (int(pr.mask[i >> 5]) >> (i & 31)) & 1
        good = ch.upper() in code and ord(e["qual"][i]) >= 33 + e["min_qual"]
        assert m == (0 if good else 1)
        if ch.upper() in code:
            assert (int(w) >> sh) & 3 == code[ch.upper()]
    # usable k-mer windows == the reference's regex segmentation
    usable = np.array([(int(pr.mask[i >> 5]) >> (i & 31)) & 1 == 0 for i in range(len(e["seq"]))])
    k = e["k"]
    win = [bool(usable[p:p + k].all()) for p in range(len(e["seq"]) - k + 1)]
    want = [any(s <= p and p + k <= t for s, t in e["fastq"]) for p in range(len(e["seq"]) - k + 1)]
    assert win == want


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "rna-bloom_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liboracle|orc_", text, re.M), f


def test_jni_shim_compiles_against_stub():
    """No JDK in the image: the shim is at least checked for syntax/type errors against a stand-in jni.h."""
    import subprocess
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "jni_stub"),
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", "rnabloom_jni.c")])
    text = open(os.path.join(ROOT, "java", "rnabloom", "gpu", "Native.java")).read()
    natives = set(re.findall(r"public static native \w+ (\w+)\(", text))
    shim = set(re.findall(r"Java_rnabloom_gpu_Native_(\w+)\(", open(os.path.join(ROOT, "jni", "rnabloom_jni.c")).read()))
    assert natives == shim, (natives ^ shim)
