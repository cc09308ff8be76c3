"""Helpers shared by the parity tests (CPU and GPU)."""
import numpy as np

from oracle import pyref as P


def np_slots(base, k, h, size):
    """numpy restatement of NTM64 + getIndex for many base hashes: (n, h) indices (NTHash.java:518-527, BloomFilter.java:108-111)."""
    base = np.asarray(base, dtype=np.int64).view(np.uint64)
    ks = np.uint64((k * P.MULTI_SEED) & P.M64)
    cols = [base >> np.uint64(1)]
    for i in range(1, h):
        t = base * (np.uint64(i) ^ ks)
        t ^= t >> np.uint64(27)
        cols.append(t >> np.uint64(1))
    return np.stack(cols, axis=1) % np.uint64(size)


def all_bases(orc, seqs, k, modes):
    """Base hashes of every k-mer window of every read under each strand mode (a superset of what was inserted)."""
    out = [orc.kmer_hashes(s, k, m)[2] for s in seqs for m in modes if len(s) >= k]
    return np.unique(np.concatenate(out)) if out else np.zeros(0, np.int64)


def counters_that_may_differ(bases, k, hc, cbf_bytes):
    """Counters of k-mers that share a counter with another distinct k-mer: there (and only there) the counting filter is order
    dependent in the reference itself, so a parallel run may legitimately differ."""
    slots = np_slots(bases, k, hc, cbf_bytes)
    uniq, cnt = np.unique(slots.reshape(-1), return_counts=True)
    shared = uniq[cnt > 1]
    touched = np.isin(slots, shared).any(axis=1)
    return set(int(x) for x in slots[touched].reshape(-1)), float(touched.mean())


def assert_cbf_close(got, want, bases, k, hc, cbf_bytes, max_frac=0.02):
    diff = np.nonzero(np.asarray(got) != np.asarray(want))[0]
    if len(diff):
        allowed, frac = counters_that_may_differ(bases, k, hc, cbf_bytes)
        assert frac < max_frac
        assert set(diff.tolist()) <= allowed, "cbf differs on counters that no other k-mer shares"
