"""Hash-sharded BloomFilterDeBruijnGraph over the GPUs of one box: one process per GPU, filters split by index range, probes routed
to the owner of their index with all-to-all exchanges (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).

This module is host plumbing only.  Every phase between two exchanges is a CUDA kernel behind the C-ABI (`rb_shard_*`,
include/rnabloom_gpu.h, kernels in csrc/rb_shard.cuh); `GpuBackend` forwards to it.  The orchestration takes the backend as a
parameter so that the exchange protocol itself (region layout, reply positions, round structure) can be exercised on CPU
with a stand-in backend that lives in tests/ -- the product has no CPU path.

The logical filters are exactly the reference's single arrays (graph/BloomFilterDeBruijnGraph.java:75-104): concatenating the ranks'
shares gives the byte array a single-GPU (or Java) run produces.
"""
import contextlib
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import binding as B
from .binding import RBError

POLICY_ADD, POLICY_COUNT_IF_PRESENT, POLICY_DBG_ONLY = 0, 1, 2


class GpuBackend:
    """rb_shard_* over device tensors (pointers are passed straight through)."""

    def __init__(self, ctx, n_ranks, rank, dbg_bits, cbf_bytes, hd, hc, k, stranded, max_kmers_per_round):
        self.ctx = ctx
        self.L = ctx.L
        h = C.c_void_p()
        ctx.check(self.L.rb_shard_create(ctx.h, n_ranks, rank, dbg_bits, cbf_bytes, hd, hc, k, int(stranded), max_kmers_per_round, C.byref(h)))
        self.h = h
        geom = (C.c_int64 * 10)()
        ctx.check(self.L.rb_shard_geometry(self.h, geom))
        (self.cap_keys, self.cap_dbg, self.cap_cbf, self.cap_lookup, self.dbg_shard, self.cbf_shard, self.local_dbg_bits,
         self.local_cbf_bytes, self.regions_per_rank, self.count_stride) = [int(x) for x in geom]
        self.device = torch.device("cuda", ctx.device)
        # kernels, torch tensor ops and the NCCL exchanges are all ordered on one (non-default) stream
        self.stream = torch.cuda.Stream(device=self.device)
        ctx.set_stream(self.stream.cuda_stream)

    def close(self):
        if self.h:
            self.ctx.check(self.L.rb_shard_destroy(self.h))
            self.h = None

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _reads(self, reads):
        # reads: (packed_ptr, mask_ptr, read_off_ptr, read_len_ptr, n_reads, uniform_len, uniform_stride) with raw device pointers
        return reads

    def route_keys(self, reads, flags, send, cnt):
        n = C.c_int64()
        self.ctx.check(self.L.rb_shard_route_keys(self.h, *reads, flags, self._p(send), self._p(cnt), C.byref(n)))
        return n.value

    def aggregate(self, recv, recv_cnt):
        self.ctx.check(self.L.rb_shard_aggregate(self.h, self._p(recv), self._p(recv_cnt)))

    def emit_dbg(self, send, cnt):
        self.ctx.check(self.L.rb_shard_emit_dbg(self.h, self._p(send), self._p(cnt)))

    def apply_dbg(self, recv, recv_cnt, reply, set_bits):
        self.ctx.check(self.L.rb_shard_apply_dbg(self.h, self._p(recv), self._p(recv_cnt), self._p(reply), int(set_bits)))

    def emit_cbf_reads(self, reply_home, policy, send, cnt):
        self.ctx.check(self.L.rb_shard_emit_cbf_reads(self.h, self._p(reply_home), policy, self._p(send), self._p(cnt)))

    def apply_cbf_read(self, recv, recv_cnt, reply):
        self.ctx.check(self.L.rb_shard_apply_cbf_read(self.h, self._p(recv), self._p(recv_cnt), self._p(reply)))

    def emit_cbf_raises(self, reply_home, policy, send, cnt):
        self.ctx.check(self.L.rb_shard_emit_cbf_raises(self.h, self._p(reply_home), policy, self._p(send), self._p(cnt)))

    def apply_cbf_raise(self, recv, recv_cnt):
        self.ctx.check(self.L.rb_shard_apply_cbf_raise(self.h, self._p(recv), self._p(recv_cnt)))

    def route_lookup(self, reads, send, cnt, fhash=None, rhash=None):
        n = C.c_int64()
        self.ctx.check(self.L.rb_shard_route_lookup(self.h, *reads, self._p(send), self._p(cnt), self._p(fhash), self._p(rhash), C.byref(n)))
        return n.value

    def apply_lookup(self, recv, recv_cnt, reply):
        self.ctx.check(self.L.rb_shard_apply_lookup(self.h, self._p(recv), self._p(recv_cnt), self._p(reply)))

    def combine_lookup(self, reply_home, counts):
        self.ctx.check(self.L.rb_shard_combine_lookup(self.h, self._p(reply_home), self._p(counts)))

    def overflow(self):
        f = C.c_int32()
        self.ctx.check(self.L.rb_shard_overflow(self.h, C.byref(f)))
        return bool(f.value)

    def local_filter(self, which):
        from .filters import BloomFilter, CountingBloomFilter
        h = C.c_void_p()
        self.ctx.check(self.L.rb_shard_filter(self.h, which, C.byref(h)))
        cls = BloomFilter if which == B.RB_DBGBF else CountingBloomFilter
        return cls(self.ctx, 0, 0, 0, _handle=h)

    def download(self, which):
        return self.local_filter(which).download()

    def popcount(self, which):
        return self.local_filter(which).getPopCount()


class ShardedGraph:
    """graph.add / graph.getKmers for reads that live on this rank, against filters sharded over all ranks."""

    def __init__(self, backend, rank, world, group=None):
        self.be, self.rank, self.world, self.group = backend, rank, world, group
        self.cap_max = max(backend.cap_keys, backend.cap_dbg, backend.cap_cbf, backend.cap_lookup)
        self.rpr = getattr(backend, "regions_per_rank", 1)   # send regions per destination rank
        cs = getattr(backend, "count_stride", 1)
        dev = backend.device
        n = world * self.rpr * self.cap_max
        self.send = torch.empty(n, dtype=torch.int64, device=dev)
        self.recv = self.send if world == 1 else torch.empty(n, dtype=torch.int64, device=dev)
        self.cnt_s = torch.zeros(world * self.rpr * cs, dtype=torch.int32, device=dev)
        self.cnt_r = self.cnt_s if world == 1 else torch.zeros(world * self.rpr * cs, dtype=torch.int32, device=dev)
        self.reply = torch.empty(n, dtype=torch.uint8, device=dev)
        self.reply_home = self.reply if world == 1 else torch.empty(n, dtype=torch.uint8, device=dev)
        self.exchanged_bytes = 0

    # ---- exchanges -------------------------------------------------------------------------------------------------------
    def _forward(self, cap):
        """send regions [world][cap] + counts -> owners."""
        if self.world == 1:
            return
        n = self.world * self.rpr * cap
        dist.all_to_all_single(self.cnt_r, self.cnt_s, group=self.group)
        dist.all_to_all_single(self.recv[:n], self.send[:n], group=self.group)
        self.exchanged_bytes += n * 8

    def _backward(self, cap):
        """reply regions travel back to where the probes came from (same offsets)."""
        if self.world == 1:
            return
        n = self.world * self.rpr * cap
        dist.all_to_all_single(self.reply_home[:n], self.reply[:n], group=self.group)
        self.exchanged_bytes += n

    # ---- one round = at most max_kmers_per_round k-mers per rank ------------------------------------------------------------
    def _on_stream(self):
        s = getattr(self.be, "stream", None)
        return torch.cuda.stream(s) if s is not None else contextlib.nullcontext()

    def add_round(self, reads, flags=0):
        with self._on_stream():
            return self._add_round(reads, flags)

    def count_round(self, reads, counts, fhash=None, rhash=None):
        with self._on_stream():
            return self._count_round(reads, counts, fhash, rhash)

    def _add_round(self, reads, flags=0):
        be = self.be
        policy = POLICY_DBG_ONLY if flags & B.DBG_ONLY else POLICY_COUNT_IF_PRESENT if flags & B.ADD_COUNT_IF_PRESENT else POLICY_ADD
        n = be.route_keys(reads, flags, self.send, self.cnt_s)
        self._forward(be.cap_keys)
        be.aggregate(self.recv, self.cnt_r)
        be.emit_dbg(self.send, self.cnt_s)
        self._forward(be.cap_dbg)
        be.apply_dbg(self.recv, self.cnt_r, self.reply, set_bits=policy != POLICY_COUNT_IF_PRESENT)
        if policy == POLICY_DBG_ONLY:
            return n
        self._backward(be.cap_dbg)
        be.emit_cbf_reads(self.reply_home, policy, self.send, self.cnt_s)
        self._forward(be.cap_cbf)
        be.apply_cbf_read(self.recv, self.cnt_r, self.reply)
        self._backward(be.cap_cbf)
        be.emit_cbf_raises(self.reply_home, policy, self.send, self.cnt_s)
        self._forward(be.cap_cbf)
        be.apply_cbf_raise(self.recv, self.cnt_r)
        return n

    def _count_round(self, reads, counts, fhash=None, rhash=None):
        be = self.be
        n = be.route_lookup(reads, self.send, self.cnt_s, fhash, rhash)
        self._forward(be.cap_lookup)
        be.apply_lookup(self.recv, self.cnt_r, self.reply)
        self._backward(be.cap_lookup)
        be.combine_lookup(self.reply_home, counts)
        return n

    def check_overflow(self):
        """A send region overflowed (pathologically skewed hashes): results of the round are incomplete -> loud failure."""
        flag = torch.tensor([1 if self.be.overflow() else 0], dtype=torch.int32, device=self.be.device)
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()):
            raise RBError(-6, "sharded exchange: a send region overflowed; lower max_kmers_per_round")

    # ---- whole logical arrays (tests, save) --------------------------------------------------------------------------------
    def gather_filter(self, which, total_bytes):
        """Concatenate the ranks' shares -> the reference's single byte array (valid on every rank)."""
        local = np.ascontiguousarray(self.be.download(which))
        if self.world == 1:
            return local[:total_bytes]
        share = (self.be.dbg_shard // 8) if which == B.RB_DBGBF else self.be.cbf_shard
        buf = torch.zeros(share, dtype=torch.uint8)
        buf[: min(len(local), share)] = torch.from_numpy(local[:share].copy())
        out = [torch.zeros(share, dtype=torch.uint8) for _ in range(self.world)]
        dev = self.be.device
        if dev.type == "cuda":   # NCCL groups only move device tensors
            outd = [o.to(dev) for o in out]
            dist.all_gather(outd, buf.to(dev), group=self.group)
            out = [o.cpu() for o in outd]
        else:
            dist.all_gather(out, buf, group=self.group)
        return torch.cat(out).numpy()[:total_bytes]


# ==== second generation: the hash-sharded graph on the sliced engine (rb_sshard_*, csrc/rb_sshard_host.inl) ==========================
class SlicedBackend:
    """rb_sshard_* over tensors whose storage the library reads and writes through raw pointers (device tensors on a GPU; the CPU
    tests run the same calls against the host emulation of the kernels, where "device memory" is host memory)."""

    def __init__(self, ctx, n_ranks, rank, dbg_bits, cbf_bytes, hd, hc, k, stranded, max_kmers_per_round, device=None):
        self.ctx, self.L = ctx, ctx.L
        h = C.c_void_p()
        ctx.check(self.L.rb_sshard_create(ctx.h, n_ranks, rank, dbg_bits, cbf_bytes, hd, hc, k, int(stranded), max_kmers_per_round, C.byref(h)))
        self.h = h
        geom = (C.c_int64 * 11)()
        ctx.check(self.L.rb_sshard_geometry(self.h, geom))
        (self.probe_regions, self.probe_cap, self.key_ranges, self.key_cap, self.raise_regions, self.raise_cap, self.dbg_share_bits,
         self.cbf_share_bytes, self.local_dbg_bits, self.local_cbf_bytes, self.spill) = [int(x) for x in geom]
        self.device = device if device is not None else torch.device("cuda", ctx.device)
        self.stream = None
        if self.device.type == "cuda":   # kernels, torch tensor ops and the NCCL exchanges are all ordered on one (non-default) stream
            self.stream = torch.cuda.Stream(device=self.device)
            ctx.set_stream(self.stream.cuda_stream)

    def close(self):
        if self.h:
            self.ctx.check(self.L.rb_sshard_destroy(self.h))
            self.h = None

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def route_lookup(self, reads, send, cnt, fhash=None, rhash=None):
        n = C.c_int64()
        self.ctx.check(self.L.rb_sshard_route_lookup(self.h, *reads, self._p(send), self._p(cnt), self._p(fhash), self._p(rhash), C.byref(n)))
        return n.value

    def route_keys(self, reads, flags, send, cnt):
        n = C.c_int64()
        self.ctx.check(self.L.rb_sshard_route_keys(self.h, *reads, flags, self._p(send), self._p(cnt), C.byref(n)))
        return n.value

    def apply(self, recv, recv_cnt, recv_ans, set_bits):
        self.ctx.check(self.L.rb_sshard_apply(self.h, self._p(recv), self._p(recv_cnt), self._p(recv_ans), int(set_bits)))

    def combine_lookup(self, home_ans, counts):
        self.ctx.check(self.L.rb_sshard_combine_lookup(self.h, self._p(home_ans), self._p(counts)))

    def dedup(self, recv, recv_cnt):
        self.ctx.check(self.L.rb_sshard_dedup(self.h, self._p(recv), self._p(recv_cnt)))

    def emit_probes(self, with_cbf, send, cnt):
        self.ctx.check(self.L.rb_sshard_emit_probes(self.h, int(with_cbf), self._p(send), self._p(cnt)))

    def combine_insert(self, home_ans, policy, send, cnt):
        self.ctx.check(self.L.rb_sshard_combine_insert(self.h, self._p(home_ans), policy, self._p(send), self._p(cnt)))

    def apply_raises(self, recv, recv_cnt):
        self.ctx.check(self.L.rb_sshard_apply_raises(self.h, self._p(recv), self._p(recv_cnt)))

    def overflow(self):
        f = C.c_int32()
        self.ctx.check(self.L.rb_sshard_overflow(self.h, C.byref(f)))
        return bool(f.value)

    def share(self, which):
        """This rank's share of a filter as a (non-owning) BloomFilter / CountingBloomFilter."""
        from .filters import BloomFilter, CountingBloomFilter
        h = C.c_void_p()
        self.ctx.check(self.L.rb_sshard_filter(self.h, which, C.byref(h)))
        cls = BloomFilter if which == B.RB_DBGBF else CountingBloomFilter
        return cls(self.ctx, 0, 0, 0, _handle=h)

    def download(self, which):
        return self.share(which).download()


class SlicedShardedGraph:
    """graph.add / graph.getKmers for reads that live on this rank, against filters sharded over all ranks: the tile sort of the sliced
    engine routes (regions of one destination are contiguous), every exchange is an equal-split all-to-all of whole regions plus
    their counts, answers come back as one byte per probe at the probe's own position."""

    def __init__(self, backend, rank, world, group=None):
        self.be, self.rank, self.world, self.group = backend, rank, world, group
        be, dev = backend, backend.device
        self.n_probe = world * be.probe_regions * be.probe_cap
        self.n_key = world * be.key_ranges * be.key_cap
        self.n_raise = world * be.raise_regions * be.raise_cap

        def buf(n, dtype):
            return torch.empty(n + be.spill, dtype=dtype, device=dev)

        self.send32 = buf(max(self.n_probe, self.n_raise), torch.int32)      # probes, then raises
        self.recv32 = self.send32 if world == 1 else buf(max(self.n_probe, self.n_raise), torch.int32)
        self.send64 = buf(self.n_key, torch.int64)
        self.recv64 = self.send64 if world == 1 else buf(self.n_key, torch.int64)
        self.ans = buf(self.n_probe, torch.uint8)                            # written by the owner at the probes' positions
        self.home_ans = self.ans if world == 1 else buf(self.n_probe, torch.uint8)
        n_cnt = world * max(be.probe_regions, be.key_ranges, be.raise_regions)
        self.cnt_s = torch.zeros(n_cnt, dtype=torch.int32, device=dev)
        self.cnt_r = self.cnt_s if world == 1 else torch.zeros(n_cnt, dtype=torch.int32, device=dev)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.exchanged_bytes = 0

    def _a2a(self, dst, src, n):
        if self.world == 1:
            return
        dist.all_to_all_single(dst[:n], src[:n], group=self.group)
        self.exchanged_bytes += n * src.element_size()

    def _forward(self, send, recv, n, n_regions):
        self._a2a(self.cnt_r, self.cnt_s, self.world * n_regions)
        self._a2a(recv, send, n)

    def _agree_no_overflow(self, what):
        """Every rank learns whether any region overflowed anywhere; nothing has been modified yet when this is called."""
        self.flag[0] = 1 if self.be.overflow() else 0
        if self.world > 1:
            dist.all_reduce(self.flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(self.flag.item()):
            raise RBError(-6, "sharded graph: a region overflowed while routing %s (skewed hashes); lower max_kmers_per_round" % what)

    def _on_stream(self):
        return torch.cuda.stream(self.be.stream) if self.be.stream is not None else contextlib.nullcontext()

    def add_round(self, reads, flags=0):
        be = self.be
        policy = POLICY_DBG_ONLY if flags & B.DBG_ONLY else POLICY_COUNT_IF_PRESENT if flags & B.ADD_COUNT_IF_PRESENT else POLICY_ADD
        with self._on_stream():
            n = be.route_keys(reads, flags, self.send64, self.cnt_s)
            self._forward(self.send64, self.recv64, self.n_key, be.key_ranges)
            be.dedup(self.recv64, self.cnt_r)
            be.emit_probes(policy != POLICY_DBG_ONLY, self.send32, self.cnt_s)
            self._agree_no_overflow("keys / probes")
            self._forward(self.send32, self.recv32, self.n_probe, be.probe_regions)
            be.apply(self.recv32, self.cnt_r, self.ans, policy != POLICY_COUNT_IF_PRESENT)
            if policy == POLICY_DBG_ONLY:
                return n
            self._a2a(self.home_ans, self.ans, self.n_probe)
            be.combine_insert(self.home_ans, policy, self.send32, self.cnt_s)
            self._forward(self.send32, self.recv32, self.n_raise, be.raise_regions)
            be.apply_raises(self.recv32, self.cnt_r)
        return n

    def count_round(self, reads, counts, fhash=None, rhash=None):
        be = self.be
        with self._on_stream():
            n = be.route_lookup(reads, self.send32, self.cnt_s, fhash, rhash)
            self._agree_no_overflow("look-up probes")
            self._forward(self.send32, self.recv32, self.n_probe, be.probe_regions)
            be.apply(self.recv32, self.cnt_r, self.ans, False)
            self._a2a(self.home_ans, self.ans, self.n_probe)
            be.combine_lookup(self.home_ans, counts)
        return n

    def check_overflow(self):
        """Raise regions can only overflow after filters were modified: loud failure."""
        self.flag[0] = 1 if self.be.overflow() else 0
        if self.world > 1:
            dist.all_reduce(self.flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(self.flag.item()):
            raise RBError(-6, "sharded graph: a raise region overflowed; lower max_kmers_per_round")

    def close(self):
        self.be.close()

    def clear(self):
        """graph.clearDbgbf + clearCbf (graph :211-230) on this rank's shares."""
        self.be.share(B.RB_DBGBF).empty()
        self.be.share(B.RB_CBF).empty()

    def popcount(self, which):
        """Set bits (dbgbf) / non-zero counters (cbf) of this rank's share; the caller sums over the ranks."""
        return self.be.share(which).getPopCount()

    def gather_filter(self, which, total_bytes):
        """Concatenate the ranks' shares -> the reference's single byte array (valid on every rank)."""
        local = np.ascontiguousarray(self.be.download(which))
        share = (self.be.dbg_share_bits // 8) if which == B.RB_DBGBF else self.be.cbf_share_bytes
        if self.world == 1:
            return local[:total_bytes]
        buf = torch.zeros(share, dtype=torch.uint8)
        buf[: min(len(local), share)] = torch.from_numpy(local[:share].copy())
        out = [torch.zeros(share, dtype=torch.uint8) for _ in range(self.world)]
        dev = self.be.device
        if dev.type == "cuda":   # NCCL groups only move device tensors
            outd = [o.to(dev) for o in out]
            dist.all_gather(outd, buf.to(dev), group=self.group)
            out = [o.cpu() for o in outd]
        else:
            dist.all_gather(out, buf, group=self.group)
        return torch.cat(out).numpy()[:total_bytes]
