"""CPU stand-in for the rb_shard_* phase kernels (TEST DOUBLE, lives in tests/ only).

It lets the exchange protocol of rna-bloom_b200/sharded.py (region layout, reply positions, round structure, all-to-all pairing) run
under gloo on a machine without a GPU.  Each phase is restated in numpy from the CUDA kernels' documented semantics
(csrc/rb_shard.cuh); hashing comes from the oracle.  The product never imports this file.
"""
import numpy as np
import torch

from oracle.binding import MODE_CANON, MODE_FWD, MODE_RC
from parity_util import np_slots

GOLD = np.uint64(0x9E3779B97F4A7C15)


def home_of(keys, n_ranks):
    m = (keys.view(np.uint64) * GOLD) >> np.uint64(32)
    return ((m * np.uint64(n_ranks)) >> np.uint64(32)).astype(np.int64)


def region_cap(total, n_ranks):
    return int(total / n_ranks * 1.06) + 4096


class CpuShardBackend:
    def __init__(self, orc, n_ranks, rank, dbg_bits, cbf_bytes, hd, hc, k, stranded, max_kmers):
        self.orc, self.G, self.rank = orc, n_ranks, rank
        self.dbg_bits, self.cbf_bytes, self.hd, self.hc, self.k, self.stranded = dbg_bits, cbf_bytes, hd, hc, k, stranded
        self.dbg_shard = -(-(-(-dbg_bits // n_ranks)) // 1024) * 1024
        self.cbf_shard = -(-(-(-cbf_bytes // n_ranks)) // 4) * 4
        self.local_dbg_bits = max(0, min(self.dbg_shard, dbg_bits - self.dbg_shard * rank))
        self.local_cbf_bytes = max(0, min(self.cbf_shard, cbf_bytes - self.cbf_shard * rank))
        recv_keys = max_kmers + max_kmers // 4 + 4096
        self.cap_keys, self.cap_dbg = region_cap(recv_keys, n_ranks), region_cap(recv_keys * hd, n_ranks)
        self.cap_cbf, self.cap_lookup = region_cap(recv_keys * hc, n_ranks), region_cap(max_kmers * (hd + hc), n_ranks)
        self.dbg = np.zeros((max(self.local_dbg_bits, 32) + 7) // 8, dtype=np.uint8)
        self.cbf = np.zeros(max(self.local_cbf_bytes, 4), dtype=np.uint8)
        self.device = torch.device("cpu")
        self._overflow = False

    # ---- helpers ---------------------------------------------------------------------------------------------------------
    def _push(self, send, cnt, dest, values, cap):
        """Append values[i] to region dest[i]; returns positions (-1 on overflow)."""
        s, c = send.numpy(), cnt.numpy()
        pos = np.full(len(values), -1, dtype=np.int64)
        for i in range(len(values)):
            d = int(dest[i])
            if c[d] >= cap:
                self._overflow = True
                continue
            pos[i] = d * cap + c[d]
            s[pos[i]] = values[i]
            c[d] += 1
        return pos

    def _valid(self, recv, recv_cnt, cap):
        r, c = recv.numpy(), recv_cnt.numpy()
        return np.concatenate([np.arange(src * cap, src * cap + int(c[src])) for src in range(self.G)]) if self.G else np.zeros(0, np.int64), r

    def _bases(self, seqs, mode):
        out, ok = [], []
        for s in seqs:
            b = np.frombuffer(s.encode() if isinstance(s, str) else bytes(s), dtype=np.uint8)
            if len(b) < self.k:
                continue
            _, _, base = self.orc.kmer_hashes(b, self.k, mode)
            good = np.isin(b, np.frombuffer(b"ACGTUacgtu", dtype=np.uint8))
            bad = np.convolve(~good, np.ones(self.k, dtype=np.int64), mode="valid")
            out.append(base), ok.append(bad == 0)
        if not out:
            return np.zeros(0, np.int64), np.zeros(0, bool)
        return np.concatenate(out), np.concatenate(ok)

    # ---- insert phases -----------------------------------------------------------------------------------------------------
    def route_keys(self, seqs, flags, send, cnt):
        cnt.zero_()
        mode = MODE_CANON if not self.stranded else (MODE_RC if flags & 1 else MODE_FWD)
        base, ok = self._bases(seqs, mode)
        keys = base[ok]
        self._push(send, cnt, home_of(keys, self.G), keys, self.cap_keys)
        return len(base)

    def aggregate(self, recv, recv_cnt):
        idx, r = self._valid(recv, recv_cnt, self.cap_keys)
        self.keys, self.mult = np.unique(r[idx], return_counts=True)

    def emit_dbg(self, send, cnt):
        cnt.zero_()
        g = np_slots(self.keys, self.k, self.hd, self.dbg_bits).astype(np.int64).reshape(-1)
        dest = g // self.dbg_shard
        self.pos_dbg = self._push(send, cnt, dest, g - dest * self.dbg_shard, self.cap_dbg).reshape(-1, self.hd)

    def apply_dbg(self, recv, recv_cnt, reply, set_bits):
        idx, r = self._valid(recv, recv_cnt, self.cap_dbg)
        out = reply.numpy()
        for p in idx:
            i = int(r[p])
            old = (self.dbg[i >> 3] >> (i & 7)) & 1
            if set_bits:
                self.dbg[i >> 3] |= 1 << (i & 7)
            out[p] = old

    def emit_cbf_reads(self, reply_home, policy, send, cnt):
        cnt.zero_()
        rep = reply_home.numpy()
        present = np.array([all(p >= 0 and rep[p] for p in row) for row in self.pos_dbg], dtype=bool) if len(self.keys) else np.zeros(0, bool)
        self.inc = np.where(present, self.mult, 0) if policy == 1 else self.mult - 1 + present
        live = np.nonzero(self.inc > 0)[0]
        g = np_slots(self.keys[live], self.k, self.hc, self.cbf_bytes).astype(np.int64).reshape(-1)
        dest = g // self.cbf_shard
        self.pos_cbf = np.full((len(self.keys), self.hc), -1, dtype=np.int64)
        self.pos_cbf[live] = self._push(send, cnt, dest, g - dest * self.cbf_shard, self.cap_cbf).reshape(-1, self.hc)

    def apply_cbf_read(self, recv, recv_cnt, reply):
        idx, r = self._valid(recv, recv_cnt, self.cap_cbf)
        out = reply.numpy()
        out[idx] = self.cbf[r[idx]]

    def emit_cbf_raises(self, reply_home, policy, send, cnt):
        cnt.zero_()
        rep = reply_home.numpy()
        vals, dests = [], []
        for i in np.nonzero(self.inc > 0)[0]:
            g = np_slots(self.keys[i:i + 1], self.k, self.hc, self.cbf_bytes).astype(np.int64)[0]
            v0 = [int(rep[p]) & 0x7F for p in self.pos_cbf[i]]
            v = list(v0)
            n = int(self.inc[i])
            if policy == 1 and min(v) == 0:
                n = 0
            for _ in range(n):
                mn = min(v)
                u = self.orc.lib.orc_minifloat_increment(mn)
                v = [u if x == mn else x for x in v]
            seen = set()
            for h in range(self.hc):
                if int(g[h]) in seen:
                    continue
                seen.add(int(g[h]))
                if v[h] > v0[h]:
                    d = int(g[h]) // self.cbf_shard
                    dests.append(d), vals.append((int(g[h]) - d * self.cbf_shard) | (v[h] << 56))
        self._push(send, cnt, np.array(dests, dtype=np.int64), np.array(vals, dtype=np.int64), self.cap_cbf)

    def apply_cbf_raise(self, recv, recv_cnt):
        idx, r = self._valid(recv, recv_cnt, self.cap_cbf)
        for p in idx:
            rec = int(r[p])
            i, v = rec & ((1 << 56) - 1), (rec >> 56) & 0xFF
            self.cbf[i] = max(int(self.cbf[i]), v)

    # ---- lookup phases -------------------------------------------------------------------------------------------------------
    def route_lookup(self, seqs, send, cnt, fhash=None, rhash=None):
        cnt.zero_()
        base, ok = self._bases(seqs, MODE_FWD if self.stranded else MODE_CANON)
        H = self.hd + self.hc
        self.pos_lookup = np.full((len(base), H), -2, dtype=np.int64)
        keys = base[ok]
        gd = np_slots(keys, self.k, self.hd, self.dbg_bits).astype(np.int64)
        gc = np_slots(keys, self.k, self.hc, self.cbf_bytes).astype(np.int64)
        dd, dc = gd // self.dbg_shard, gc // self.cbf_shard
        rec = np.concatenate([gd - dd * self.dbg_shard, (gc - dc * self.cbf_shard) | np.int64(-2 ** 63)], axis=1)
        dest = np.concatenate([dd, dc], axis=1)
        self.pos_lookup[ok] = self._push(send, cnt, dest.reshape(-1), rec.reshape(-1), self.cap_lookup).reshape(-1, H)
        return len(base)

    def apply_lookup(self, recv, recv_cnt, reply):
        idx, r = self._valid(recv, recv_cnt, self.cap_lookup)
        out = reply.numpy()
        for p in idx:
            rec = int(r[p])
            i = rec & ((1 << 63) - 1)
            out[p] = self.cbf[i] if rec < 0 else (self.dbg[i >> 3] >> (i & 7)) & 1

    def combine_lookup(self, reply_home, counts):
        rep, out = reply_home.numpy(), counts.numpy()
        for j, row in enumerate(self.pos_lookup):
            c = 0.0
            if row[0] != -2 and all(p >= 0 and rep[p] for p in row[: self.hd]):
                mn = min(int(np.int8(rep[p])) if p >= 0 else 0 for p in row[self.hd:])
                c = self.orc.lib.orc_minifloat_to_float(mn) + 1.0
            out[j] = c

    def overflow(self):
        f, self._overflow = self._overflow, False
        return f

    def download(self, which):
        return self.dbg if which == 0 else self.cbf

    def close(self):
        pass
