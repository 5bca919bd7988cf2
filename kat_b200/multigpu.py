"""Multi-GPU counting: one process per GPU, reads sharded by record, k-mers
hash-partitioned to owner ranks.

A count histogram is not additive over read shards (a k-mer seen 3x on GPU0 and
2x on GPU1 belongs in bin 5), so before any reduction every distinct k-mer must
live on exactly one rank:

  route    katgpu_route_device: extract + canonicalise + hash the local base
           stream, bucket the 64-bit hashed keys by owner = f(hash) % world
  exchange all-to-all of the buckets over NVLink (NCCL through torch.distributed)
  insert   katgpu_insert_hashed_device on the owner
  reduce   local hist/gcp walk, then ONE all-reduce(sum) of the bins / matrix

The orchestration is engine-agnostic so that the host logic can be exercised
with gloo on CPU (tests/test_multigpu_gloo.py passes a numpy engine); the
product engine below drives libkatgpu and has no CPU path."""
import math

import numpy as np

from . import binding as B
from .routing import KMERS_SLACK


class GpuEngine:
    """libkatgpu + torch tensors for the exchange buffers."""

    def __init__(self, ctx, k, canonical, slots):
        import torch
        self.torch = torch
        self.ctx, self.k, self.canonical, self.slots = ctx, k, canonical, slots
        self.table = None
        self.device = torch.device("cuda", torch.cuda.current_device())

    def begin(self, expected_kmers=0):
        self.table = self.ctx.count_begin(self.k, self.canonical, self.slots, True)
        if expected_kmers:
            self.table.reserve(expected_kmers)

    def alloc_keys(self, n):
        return self.torch.empty(max(n, 1), dtype=self.torch.int64, device=self.device)

    def route(self, bases_ptr, n_bytes, out_tensors, cap):
        return self.table.route_device(bases_ptr, n_bytes, [t.data_ptr() for t in out_tensors], cap)

    def insert(self, keys_tensor, n):
        if n:
            self.table.insert_hashed_device(keys_tensor.data_ptr(), n)

    def sync_exchange(self):
        self.torch.cuda.current_stream().synchronize()

    def end(self):
        self.table.end()
        return self.table.n_distinct, self.table.n_total

    def hist(self, low, high, inc):
        return self.table.hist(low, high, inc)[1]

    def gcp(self, scale, bins):
        return self.table.gcp(scale, bins)

    def to_tensor(self, arr):
        return self.torch.from_numpy(arr.astype(np.int64)).to(self.device)

    def from_tensor(self, t):
        return t.cpu().numpy().astype(np.uint64)

    def free(self):
        if self.table is not None:
            self.table.free()
            self.table = None


class DistributedCounter:
    def __init__(self, ctx, dist, k, canonical=True, slots=1 << 31, batch_bytes=1 << 29, engine=None):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.k = k
        self.batch_bytes = batch_bytes
        self.engine = engine if engine is not None else GpuEngine(ctx, k, canonical, slots)
        self._send = self._recv = None
        self._cap = 0

    def _buffers(self, cap):
        """Two send and two receive bucket sets: batch b+1 is routed and batch b-1 is
        inserted while batch b is on the wire."""
        if cap > self._cap:
            self._send = [[self.engine.alloc_keys(cap) for _ in range(self.world)] for _ in range(2)]
            self._recv = [[self.engine.alloc_keys(cap) for _ in range(self.world)] for _ in range(2)]
            self._cap = cap
        return self._send, self._recv

    def count(self, bases_ptr, n_bytes, record_len=None):
        """Route / exchange / insert the local base stream in batches, software-pipelined:
        route(b+1) and insert(b-1) overlap the exchange of batch b.  Batches end on record
        boundaries when record_len is given (no k-mer is cut)."""
        e, dist, W = self.engine, self.dist, self.world
        # with a uniform owner hash every rank receives about what it routes
        e.begin(n_bytes)
        step = self.batch_bytes
        if record_len:  # whole records per batch, and 16-byte aligned batch starts (TMA path)
            unit = record_len * 16
            step = max(unit, step // unit * unit)
        # every rank must take part in the same number of collectives
        nb_local = max(1, math.ceil(n_bytes / step))
        nbt = e.to_tensor(np.array([nb_local]))
        dist.all_reduce(nbt, op=dist.ReduceOp.MAX)
        n_batches = int(e.from_tensor(nbt)[0])
        cap = int(step / W * 1.25) + KMERS_SLACK
        send, recv = self._buffers(cap)

        def route(b):
            off = min(n_bytes, b * step)
            n = max(0, min(step, n_bytes - off))
            counts = e.route(bases_ptr + off, n, send[b & 1], cap) if n else [0] * W
            sc = e.to_tensor(np.array(counts))
            rc = e.to_tensor(np.zeros(W))
            dist.all_to_all_single(rc, sc)
            rcounts = [int(x) for x in e.from_tensor(rc)]
            if max(rcounts) > cap:
                raise B.KatGpuError(-1, "receive bucket too small for this batch")
            return counts, rcounts

        cnt = route(0)
        works = self._exchange_start(send[0], cnt[0], recv[0], cnt[1])
        for b in range(n_batches):
            nxt = route(b + 1) if b + 1 < n_batches else None  # overlaps exchange b
            for w in works:
                w.wait()
            e.sync_exchange()
            if nxt is not None:
                works = self._exchange_start(send[(b + 1) & 1], nxt[0], recv[(b + 1) & 1], nxt[1])
            for r in range(W):  # overlaps exchange b+1
                e.insert(recv[b & 1][r], cnt[1][r])
            cnt = nxt
        return e.end()

    def _exchange_start(self, send, counts, recv, rcounts):
        """Ragged all-to-all as grouped point-to-point (what NCCL's all-to-all is
        underneath); works on NCCL and gloo alike.  Returns the pending work handles."""
        dist, W, me = self.dist, self.world, self.rank
        ops = []
        for d in range(1, W):
            to, frm = (me + d) % W, (me - d) % W
            if counts[to]:
                ops.append(dist.P2POp(dist.isend, send[to][: counts[to]], to))
            if rcounts[frm]:
                ops.append(dist.P2POp(dist.irecv, recv[frm][: rcounts[frm]], frm))
        if rcounts[me]:
            recv[me][: rcounts[me]].copy_(send[me][: counts[me]])
        return dist.batch_isend_irecv(ops) if ops else []

    def _allreduce_u64(self, arr):
        t = self.engine.to_tensor(arr.reshape(-1))
        self.dist.all_reduce(t)
        return self.engine.from_tensor(t).reshape(arr.shape)

    def hist_job(self, bases_ptr, n_bytes, low=1, high=10000, inc=1, record_len=None):
        nd, nt = self.count(bases_ptr, n_bytes, record_len)
        bins = self._allreduce_u64(self.engine.hist(low, high, inc))
        self.engine.free()
        return bins, nt, nd

    def gcp_job(self, bases_ptr, n_bytes, scale=1.0, cvg_bins=1000, record_len=None):
        nd, nt = self.count(bases_ptr, n_bytes, record_len)
        mx = self._allreduce_u64(self.engine.gcp(scale, cvg_bins))
        self.engine.free()
        return mx, nt, nd
