"""Where the time of the ASCII entry points goes: wall time of each call + the library's kernel spans (GPU box only)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import rnabloom_b200 as rb
from rnabloom_b200.filters import _ptr

K, L, STRIDE = 25, 150, 160
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = rb.Context(0)
g = rb.BloomFilterDeBruijnGraph(ctx, 1 << 36, 1 << 33, 64, 3, 3, 1, K, False, False)
w = n * STRIDE // 32
tmp = ctx.dev_alloc(w * 8 + 64)
pk = np.zeros(w, dtype=np.uint64)
ctx.synth_reads_dev(11, 3_000_000_000, 0, n, L, 5000, STRIDE, tmp)
ctx.sync()
ctx.d2h(pk, tmp)
codes = ((pk[:, None] >> (2 * np.arange(32, dtype=np.uint64))[None, :]) & np.uint64(3)).astype(np.uint8).reshape(n, STRIDE)[:, :L]
for pinned in (False, True):
    bases = ctx.host_alloc(n * L, np.uint8) if pinned else np.zeros(n * L, dtype=np.uint8)
    bases[:] = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].reshape(-1)
    off = np.arange(n + 1, dtype=np.int64) * L
    counts = ctx.host_alloc(n * (L - K + 1) * 4, np.float32)
    na = C.c_int64()
    for rep in range(3):
        ctx.profile_enable(True)
        ctx.profile_read()
        t0 = time.perf_counter()
        ctx.check(ctx.L.rb_graph_add_reads_ascii(g.h, _ptr(bases), None, _ptr(off), n, 0, 0, C.byref(na)))
        t1 = time.perf_counter()
        ctx.check(ctx.L.rb_graph_count_reads_ascii(g.h, _ptr(bases), _ptr(off), n, _ptr(counts), None, None, C.byref(na)))
        t2 = time.perf_counter()
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        print("pinned=%s rep %d: add %.1f ms, count %.1f ms; kernels %.1f ms: %s" % (
            pinned, rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), sum(v[0] for v in prof.values()),
            {k_: round(v[0], 2) for k_, v in prof.items()}), flush=True)
