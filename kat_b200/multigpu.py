"""Multi-GPU jobs, one process per GPU (torchrun): the plumbing around libkatgpu's own
multi-GPU engine (kat_b200/csrc/katgpu_dist.cu, dist_kernels.cu).

A count histogram is not additive over read shards (a k-mer seen 3x on GPU0 and 2x on GPU1
belongs in bin 5), so every distinct k-mer is counted by exactly one owner rank.  The engine
does that inside the library: reads are sharded by record, each rank feeds its shard, the
extraction kernel stores every hashed key straight into the owner's HBM over NVLink, the owner
inserts, and hist / gcp / comp end with ONE ncclAllReduce of the result.  What is left for Python:

  shard_records()        which records a rank feeds (contiguous, balanced)
  broadcast_unique_id()  rank 0's 128-byte job id to every rank through torch.distributed
  join_job()             ctx.dist_init(...) with it
  DistributedJob         `kat hist|gcp|comp` as collective calls, mirroring the C++ front-end

torch.distributed is only the rendezvous here (gloo or nccl, both work) — tests/test_multigpu_gloo.py
drives the same functions with world_size 2 and 3 on CPU.  There is no CPU engine."""
import numpy as np

from . import binding as B


def shard_records(n_records, rank, world):
    """(first, count) of the contiguous block of records rank `rank` feeds: blocks differ by at most one
    record and cover [0, n_records) exactly once."""
    base, extra = divmod(n_records, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def broadcast_unique_id(dist, make_id=B.dist_unique_id, device=None):
    """Rank 0 creates the job id (katgpu_dist_unique_id -> ncclGetUniqueId), everybody receives it."""
    import torch
    rank = dist.get_rank()
    raw = make_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def join_job(ctx, dist, device=None):
    """Make `ctx` rank dist.get_rank() of a dist.get_world_size()-GPU job.  world 1: nothing to do."""
    world = dist.get_world_size()
    if world > 1:
        ctx.dist_init(dist.get_rank(), world, broadcast_unique_id(dist, device=device))
    return ctx


class DistributedJob:
    """Collective `kat` jobs on a joined context: every rank calls the same methods in the same
    order with ITS shard of the input; every rank gets the complete result."""

    def __init__(self, ctx, k, canonical=True, slots=1 << 31):
        self.ctx, self.k, self.canonical, self.slots = ctx, k, canonical, slots

    def count(self, bases_ptr, n_bytes, device=True, expected_kmers=0):
        t = self.ctx.count_begin(self.k, self.canonical, self.slots, True)
        if expected_kmers:
            t.reserve(expected_kmers)
        if device:
            t.feed_device(bases_ptr, n_bytes)
        else:
            t.feed_ptr(bases_ptr, n_bytes)
        return t.end()

    def hist_job(self, bases_ptr, n_bytes, low=1, high=10000, inc=1, device=True):
        t = self.count(bases_ptr, n_bytes, device)
        _, bins = t.hist(low, high, inc)
        nt, nd = t.n_total, t.n_distinct
        t.free()
        return bins, nt, nd

    def gcp_job(self, bases_ptr, n_bytes, scale=1.0, cvg_bins=1000, device=True):
        t = self.count(bases_ptr, n_bytes, device)
        mx = t.gcp(scale, cvg_bins)
        nt, nd = t.n_total, t.n_distinct
        t.free()
        return mx, nt, nd

    def comp_job(self, ptr1, n1, ptr2, n2, device=True, **kw):
        t1 = self.count(ptr1, n1, device)
        t2 = self.count(ptr2, n2, device)
        r = B.comp(t1, t2, **kw)
        r["n_total"] = (t1.n_total, t2.n_total)
        t1.free()
        t2.free()
        return r


def simulate_partitioned_counts(keys, world, k):
    """What the engine does, restated on the host for one rank's canonical k-mer instances `keys`
    (uint64): the hashed keys and their owner ranks (kat_b200/routing.py = table.cuh).  Used by the CPU
    tests to check that owner-partitioned counting followed by a sum reproduces the single table."""
    from .routing import mix_key, owner_of
    h = mix_key(np.asarray(keys, dtype=np.uint64), k)
    return h, owner_of(h, world)
