"""Hash-sharded BloomFilterDeBruijnGraph over the GPUs of one box: one process per GPU, filters split by index range, every probe routed
to the owner of its index (include/rnabloom_gpu.h `rb_mgraph_*`, csrc/rb_mgraph_host.inl).

This module is a thin host mirror: the round orchestration AND the exchanges live in the library (NCCL bound at run time, on the
library's stream); Python only hands over device pointers and, once, the NCCL unique id that the ranks have to share.  The logical
filters are exactly the reference's single arrays (graph/BloomFilterDeBruijnGraph.java:75-104): `gather_filter` reassembles the ranks'
shares into the byte array a single-GPU (or Java) run produces.

Transports:
  * ``nccl_id`` (default on GPUs): rank 0 calls `nccl_unique_id()`, the bytes travel over any host channel (here: torch.distributed
    broadcast), every rank passes them to `ShardedGraph`; the library then owns its own NCCL communicator.
  * ``GlooTransport``: the CPU tests run the same library orchestrator over the host emulation of the kernels ("device memory" is host
    memory there) with the two transport callbacks implemented on torch.distributed / gloo.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import binding as B
from .binding import RBError  # noqa: F401

POLICY_ADD, POLICY_COUNT_IF_PRESENT, POLICY_DBG_ONLY = 0, 1, 2

_A2A = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
_ARM = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
_AG = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class _Transport(C.Structure):
    _fields_ = [("user", C.c_void_p), ("all_to_all", _A2A), ("all_reduce_max", _ARM), ("all_gather", _AG)]


def nccl_unique_id(lib=None):
    """128 bytes from ncclGetUniqueId (call on rank 0, distribute to every rank)."""
    L = lib or B.lib()
    buf = (C.c_uint8 * 128)()
    rc = L.rb_nccl_unique_id(buf, 128)
    if rc:
        raise RBError(rc, (L.rb_last_error(None) or b"").decode())
    return bytes(buf)


def broadcast_nccl_id(device=None, group=None):
    """rank 0 creates the id, torch.distributed carries it to the other ranks (any channel would do: it is 128 bytes of host data)."""
    t = torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank(group) == 0:
        t = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).clone()
    if device is not None and device.type == "cuda":
        t = t.to(device)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().numpy().tobytes())


class GlooTransport:
    """rb_transport over torch.distributed for HOST memory (CPU tests with the emulated kernels)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)

        def view(ptr, n, dtype):
            item = torch.empty(0, dtype=dtype).element_size()
            return torch.frombuffer((C.c_uint8 * (n * item)).from_address(ptr), dtype=dtype)

        def a2a(_user, send, recv, bytes_per_rank, _stream):
            try:
                n = bytes_per_rank * self.world
                dist.all_to_all_single(view(recv, n, torch.uint8), view(send, n, torch.uint8).clone(), group=self.group)
                return 0
            except Exception:   # noqa: BLE001 -- an exception must not unwind through the C frames
                return -4

        def arm(_user, buf, n, _stream):
            try:
                dist.all_reduce(view(buf, n, torch.int32), op=dist.ReduceOp.MAX, group=self.group)
                return 0
            except Exception:   # noqa: BLE001
                return -4

        self._a2a, self._arm = _A2A(a2a), _ARM(arm)   # keep the callbacks alive
        self.struct = _Transport(None, self._a2a, self._arm, _AG())   # no all_gather: host memory has no peer-to-peer mode


class ShardedGraph:
    """graph.add / graph.getKmers for reads that live on this rank, against filters sharded over all ranks."""

    def __init__(self, ctx, n_ranks, rank, dbg_bits, cbf_bytes, hd, hc, k, stranded, max_kmers_per_round, nccl_id=None, transport=None,
                 device=None):
        self.ctx, self.L, self.rank, self.world = ctx, ctx.L, rank, n_ranks
        self.dbg_bits, self.cbf_bytes = dbg_bits, cbf_bytes
        self.device = device if device is not None else torch.device("cuda", ctx.device)
        self._transport = transport
        h = C.c_void_p()
        if transport is not None or n_ranks == 1:
            tp = C.byref(transport.struct) if transport is not None else None
            ctx.check(self.L.rb_mgraph_create(ctx.h, n_ranks, rank, tp, dbg_bits, cbf_bytes, hd, hc, k, int(stranded), max_kmers_per_round, C.byref(h)))
        else:
            if nccl_id is None:
                raise ValueError("ShardedGraph over several GPUs needs nccl_id (see broadcast_nccl_id) or a transport")
            idbuf = (C.c_uint8 * 128).from_buffer_copy(nccl_id)
            ctx.check(self.L.rb_mgraph_create_nccl(ctx.h, n_ranks, rank, idbuf, dbg_bits, cbf_bytes, hd, hc, k, int(stranded), max_kmers_per_round,
                                                   C.byref(h)))
        self.h = h
        lay = (C.c_int64 * 9)()
        ctx.check(self.L.rb_mgraph_layout(self.h, lay))
        (self.paired, self.dbg_share_bits, self.cbf_share_bytes, self.local_dbg_bits, self.local_cbf_bytes, self.chunks, self.max_kmers,
         self.insert_bytes_per_round, self.lookup_bytes_per_round) = [int(x) for x in lay]

    def close(self):
        if self.h:
            self.ctx.check(self.L.rb_mgraph_destroy(self.h))
            self.h = None

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    @property
    def peer_to_peer(self):
        return bool(self.L.rb_mgraph_peer_to_peer(self.h))

    @property
    def exchanged_bytes(self):
        x, r = C.c_int64(), C.c_int64()
        self.ctx.check(self.L.rb_mgraph_stats(self.h, C.byref(x), C.byref(r)))
        return x.value

    def add_round(self, reads, flags=0):
        """reads = (packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride) with raw device pointers."""
        n = C.c_int64()
        self.ctx.check(self.L.rb_mgraph_add_round_dev(self.h, *reads, flags, C.byref(n)))
        return n.value

    def count_round(self, reads, counts, fhash=None, rhash=None):
        n = C.c_int64()
        self.ctx.check(self.L.rb_mgraph_count_round_dev(self.h, *reads, self._p(counts), self._p(fhash), self._p(rhash), C.byref(n)))
        return n.value

    def check_overflow(self):
        """Kept for callers of the earlier interface: overflow is reported by the round calls themselves now (RBError -6)."""

    def share(self, which):
        """This rank's share of a filter as a (non-owning) BloomFilter / CountingBloomFilter."""
        from .filters import BloomFilter, CountingBloomFilter
        h = C.c_void_p()
        self.ctx.check(self.L.rb_mgraph_filter(self.h, which, C.byref(h)))
        cls = BloomFilter if which == B.RB_DBGBF else CountingBloomFilter
        return cls(self.ctx, 0, 0, 0, _handle=h)

    def clear(self):
        """graph.clearDbgbf + clearCbf (graph :211-230) on this rank's shares."""
        self.share(B.RB_DBGBF).empty()
        self.share(B.RB_CBF).empty()

    def popcount(self, which):
        """Set bits (dbgbf) / non-zero counters (cbf) of this rank's share; the caller sums over the ranks."""
        return self.share(which).getPopCount()

    def _all_gather(self, local, nbytes, group=None):
        buf = torch.zeros(nbytes, dtype=torch.uint8)
        buf[: min(len(local), nbytes)] = torch.from_numpy(np.ascontiguousarray(local[:nbytes]).copy())
        if self.world == 1:
            return [buf]
        out = [torch.zeros(nbytes, dtype=torch.uint8) for _ in range(self.world)]
        if self.device.type == "cuda":   # NCCL groups only move device tensors
            outd = [o.to(self.device) for o in out]
            dist.all_gather(outd, buf.to(self.device), group=group)
            return [o.cpu() for o in outd]
        dist.all_gather(out, buf, group=group)
        return out

    def gather_filter(self, which, total_bytes, group=None):
        """Reassemble the ranks' shares -> the reference's single byte array (valid on every rank).  Test / save path: host memory."""
        local = self.share(which).download()
        if which == B.RB_CBF or not self.paired:
            share = self.cbf_share_bytes if which == B.RB_CBF else self.dbg_share_bits // 8
            return torch.cat(self._all_gather(local, share, group)).numpy()[:total_bytes]
        # paired slices: rank r holds, for every chunk c of cbf_bytes bits, the bits [r * share, (r + 1) * share) of that chunk
        per = self.cbf_share_bytes // 8                       # bytes of one chunk piece
        parts = self._all_gather(local, per * self.chunks, group)
        arr = np.stack([p.numpy().reshape(self.chunks, per) for p in parts], axis=1)   # [chunk][rank][bytes]
        return arr.reshape(-1)[:total_bytes]
