"""util/SeqSubsampler.java on the batched primitives of librnabloom_gpu.so.

`minimizerBased` (util/SeqSubsampler.java:50-117) is a greedy loop over the sequences, longest first: a sequence is kept when too few of its
minimizers have been seen often enough before; every minimizer of every sequence goes through `CountingBloomFilter.incrementAndGet`.  The
key generator is data parallel -- `rb_minimizers` produces `MinimizerHashIterator.next()` of every window of every sequence in one launch --
the loop itself is sequential by definition (the decision for a sequence depends on the counts left by the ones before it), so it stays on the
host and feeds the counting filter one sequence at a time (`rb_cbf_increment_and_get_hashes`).
"""
import numpy as np

from .binding import MODE_CANON, MODE_FWD
from .filters import CountingBloomFilter, pack_reads


def compress_homopolymers(seq):
    """util/SeqUtils.java:1708-1730"""
    if not seq:
        return seq
    out = [seq[0]]
    for ch in seq[1:]:
        if ch != out[-1]:
            out.append(ch)
    return "".join(out)


def _increment_and_get(cbf, keys):
    """incrementAndGet of every key, in order.  A batch is linearised per key instance on the device; a key that repeats inside one sequence
    must see its own earlier increment, so the list is cut where a key repeats."""
    out = np.zeros(len(keys), dtype=np.float32)
    lo, seen = 0, set()
    for i, key in enumerate(keys):
        if key in seen:
            out[lo:i] = cbf.incrementAndGet(np.asarray(keys[lo:i], dtype=np.int64))
            lo, seen = i, set()
        seen.add(key)
    if lo < len(keys):
        out[lo:] = cbf.incrementAndGet(np.asarray(keys[lo:], dtype=np.int64))
    return out


def minimizerBased(ctx, seqs, bfSize, k, w, numHash, stranded, useHpcKmers, maxNonMatchingChainLength, minMatchingProportion, maxMultiplicity):
    """Returns the indices of the sequences the reference would write to the subsample (util/SeqSubsampler.java:50-117) and the filter."""
    texts = [compress_homopolymers(s) if useHpcKmers else s for s in seqs]
    reads = pack_reads(texts)
    mins = ctx.minimizers(reads, k, w, MODE_FWD if stranded else MODE_CANON)
    off = reads.offsets(k + w - 1)
    cbf = CountingBloomFilter(ctx, bfSize, numHash, k)
    kept = []
    for i in range(len(seqs)):
        m = mins[off[i]:off[i + 1]]
        if m.size == 0:                       # itr.start(hpc) is false: too short for one window
            kept.append(i)
            continue
        keys = [int(m[0])] + [int(b) for a, b in zip(m[:-1], m[1:]) if a != b]   # "if (mm != prev)"
        counts = _increment_and_get(cbf, keys)
        num_seen = int(counts[0] > maxMultiplicity)
        consecutive = max_consecutive = 0
        for c in counts[1:]:
            if c > maxMultiplicity:
                num_seen += 1
                consecutive = 0
            else:
                consecutive += 1
                max_consecutive = max(max_consecutive, consecutive)
        if max_consecutive > maxNonMatchingChainLength or num_seen < minMatchingProportion * len(keys):
            kept.append(i)
    return kept, cbf
