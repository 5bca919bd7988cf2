"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: record sharding, the job-id
broadcast, and the algorithm the engine implements — owner routing of every k-mer instance
(kat_b200/routing.py restates table.cuh), a ragged all-to-all, per-owner counting, one all-reduce of
the bins — against the oracle's single table.  The CUDA engine itself (route_bases_kernel + peer
stores) is covered on GPUs by tests/test_gpu_multi.py and bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
import seqgen
from kat_b200.multigpu import broadcast_unique_id, shard_records, simulate_partitioned_counts
from kat_b200.routing import mix_key, owner_of


def _worker(rank, world, port, k, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the rendezvous the GPU job uses: rank 0's 128-byte id reaches everybody unchanged
        uid = broadcast_unique_id(dist, make_id=lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        n_records, rec = 1500, 101
        full = seqgen.genome_reads(20_000, n_records, 100, seed=5)
        first, cnt = shard_records(n_records, rank, world)
        mine = full[first * rec:(first + cnt) * rec]
        # this rank's canonical k-mer instances (the oracle stands in for the extraction kernel)
        t = O.Table.from_chunks([mine.tobytes()], k, True)
        keys, counts = t.export()
        inst = np.repeat(keys, counts.astype(np.int64))
        h, own = simulate_partitioned_counts(inst, world, k)
        # ragged all-to-all of the hashed keys to their owners
        send = [h[own == r] for r in range(world)]
        sc = torch.tensor([len(s) for s in send])
        rc = torch.zeros(world, dtype=torch.int64)
        dist.all_to_all_single(rc, sc)
        recv = torch.empty(int(rc.sum()), dtype=torch.int64)
        dist.all_to_all_single(recv, torch.from_numpy(np.concatenate(send).view(np.int64).copy()),
                               output_split_sizes=[int(n) for n in rc], input_split_sizes=[int(n) for n in sc])
        got = recv.numpy().view(np.uint64)
        assert np.all(owner_of(got, world) == rank)  # every received key is mine
        uniq, cts = np.unique(got, return_counts=True)
        bins = np.zeros(10001, dtype=np.int64)
        np.add.at(bins, np.minimum(cts, 10001) - 1, 1)
        tb = torch.from_numpy(bins)
        dist.all_reduce(tb)  # the one all-reduce of the result
        tot = torch.tensor([len(got), len(uniq)])
        dist.all_reduce(tot)
        if rank == 0:
            np.savez(q, bins=tb.numpy().astype(np.uint64), tot=tot.numpy())
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


def test_shard_records_covers_everything_once():
    for n in (0, 1, 7, 1500, 66_666_667):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_records(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


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
