"""rna-bloom_b200 -- host-side mirror of RNA-Bloom's BloomFilter / CountingBloomFilter / BloomFilterDeBruijnGraph
operator surface over librnabloom_gpu.so (hand-written CUDA for sm_100a, C-ABI in include/rnabloom_gpu.h).

The directory name has a hyphen (it is the name the project asks for); import it as ``rnabloom_b200`` through the
shim module at the repo root.  There is no CPU fallback: without the CUDA library or a GPU every call raises.
"""
from .binding import (  # noqa: F401
    RBError, lib, lib_path, build_library,
    RB_BLOOM, RB_COUNTING, RB_DBGBF, RB_CBF, RB_RPKBF, RB_FPKBF,
    MODE_FWD, MODE_RC, MODE_CANON,
    REVCOMP, ADD_COUNT_IF_PRESENT, DBG_ONLY, STORE_READ_PAIRS, STORE_FRAG_PAIRS, PAIRS_EXISTING_ONLY,
)
from .filters import (  # noqa: F401
    Context, BloomFilter, CountingBloomFilter, CascadingBloomFilter, KmerHistogram, BloomFilterDeBruijnGraph, PackedReads, pack_reads, pack_uniform,
    encode_2bit_records,
)
