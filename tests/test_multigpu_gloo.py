"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: read
sharding, owner routing, all-to-all exchange with ragged counts, batching, and
the final all-reduce.  The CUDA engine is replaced by a numpy engine built on
the oracle (checker), so this runs without a GPU; the same orchestration class
drives libkatgpu on the B200 box (tests/test_gpu_multi.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
import seqgen
from kat_b200.multigpu import DistributedCounter
from kat_b200.routing import mix_key, owner_of


class NumpyEngine:
    def __init__(self, bases, k, canonical):
        self.bases, self.k, self.canonical = bases, k, canonical

    def begin(self, expected_kmers=0):
        self.h = []

    def alloc_keys(self, n):
        return torch.empty(max(n, 1), dtype=torch.int64)

    def route(self, ptr, n, outs, cap):
        t = O.Table.from_chunks([self.bases[ptr:ptr + n].tobytes()], self.k, self.canonical)
        keys, counts = t.export()
        inst = np.repeat(keys, counts.astype(np.int64))
        h = mix_key(inst, self.k)
        own = owner_of(h, len(outs))
        res = []
        for r, o in enumerate(outs):
            sel = h[own == r]
            assert len(sel) <= cap
            o[: len(sel)] = torch.from_numpy(sel.view(np.int64))
            res.append(len(sel))
        return res

    def insert(self, keys, n):
        self.h.append(keys[:n].numpy().view(np.uint64).copy())

    def sync_exchange(self):
        pass

    def end(self):
        allh = np.concatenate(self.h) if self.h else np.zeros(0, np.uint64)
        self.uniq, self.counts = np.unique(allh, return_counts=True)
        return len(self.uniq), len(allh)

    def hist(self, low, high, inc):
        assert (low, high, inc) == (1, 10000, 1)
        bins = np.zeros(10001, dtype=np.uint64)
        np.add.at(bins, np.minimum(self.counts, 10001) - 1, 1)
        return bins

    def to_tensor(self, arr):
        return torch.from_numpy(np.asarray(arr).astype(np.int64))

    def from_tensor(self, t):
        return t.numpy().astype(np.uint64)

    def free(self):
        pass


def _worker(rank, world, port, k, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = seqgen.genome_reads(20_000, 1500, 100, seed=5)
        rec = 101
        per = (1500 + world - 1) // world
        mine = full[rank * per * rec: min(1500, (rank + 1) * per) * rec]
        eng = NumpyEngine(mine, k, True)
        dc = DistributedCounter(None, dist, k, True, 0, batch_bytes=rec * 16 * 9, engine=eng)
        bins, nt, nd = dc.hist_job(0, len(mine), record_len=rec)
        tot = torch.tensor([nt, nd])
        dist.all_reduce(tot)
        if rank == 0:
            np.savez(q, bins=bins, tot=tot.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_hist_equals_single_table(world, tmp_path):
    k = 21
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(world, _free_port(), k, out), nprocs=world, join=True)  # raises if a worker died
    r = np.load(out)
    bins, n_total, n_distinct = r["bins"], int(r["tot"][0]), int(r["tot"][1])
    full = seqgen.genome_reads(20_000, 1500, 100, seed=5)
    ot = O.Table.from_chunks([full.tobytes()], k, True)
    assert n_total == ot.total and n_distinct == ot.distinct
    assert np.array_equal(bins, ot.hist()[1])


def test_routing_restatement_properties():
    rng = np.random.default_rng(0)
    for k in (5, 21, 27, 32):
        b = 2 * k
        x = rng.integers(0, 1 << min(b, 63), 5000, dtype=np.uint64)
        h = mix_key(x, k)
        assert len(np.unique(h)) == len(np.unique(x))  # bijection: no collisions
        if b < 64:
            assert int(h.max()) < (1 << b)
        for w in (1, 2, 3, 8):
            o = owner_of(h, w)
            assert o.min() >= 0 and o.max() < w
            if w > 1 and k > 5:
                assert np.bincount(o, minlength=w).min() > len(h) / w * 0.8  # balanced
