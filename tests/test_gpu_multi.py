"""N>1 on real GPUs (skipped on a 1-GPU box): libkatgpu's multi-GPU engine must reproduce the
single-table oracle result bit for bit — hist, gcp and comp — (a) with one process per GPU under
torchrun (peer memory through CUDA IPC, job id through torch.distributed) from device-resident and
from host streams, with ragged and empty shards and several route rounds, and (b) through the C++
front-end, `kat --gpus 2`, one process driving both GPUs, output files byte-identical to the
reference's."""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import golden_util as G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = os.path.join(ROOT, "kat_b200", "bin", "kat")

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import kat_b200
from kat_b200 import binding as B
from kat_b200.multigpu import DistributedJob, join_job, shard_records
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
ctx = join_job(kat_b200.Context(lr), dist)
assert (ctx.rank, ctx.n_ranks) == (rank, world)
n_reads, L, k = 60000, 150, 27
# uneven shards: the last rank gets only a handful of reads (fewer route rounds than the others)
first, cnt = (0, n_reads - 7) if rank == 0 else (n_reads - 7, 7) if rank == world - 1 else (n_reads - 7, 0)
if world == 2 and %(even)d:
    first, cnt = shard_records(n_reads, rank, world)
a = B.synth_reads_host(300_000, 7, 8, first, cnt, L)
b = B.synth_reads_host(300_000, 7, 9, first // 2, cnt // 2, L, 0.02, 0.001)
job = DistributedJob(ctx, k, True, 1 << 22)
res = {}
for mode in ("device", "host"):
    if mode == "device":
        d = ctx.device_alloc(max(a.size, 16)); ctx.h2d(d, a) if a.size else None
        bins, nt, nd = job.hist_job(d, a.size)
        mx, _, _ = job.gcp_job(d, a.size, 0.5, 300)
        d2 = ctx.device_alloc(max(b.size, 16)); ctx.h2d(d2, b) if b.size else None
        cmp_ = job.comp_job(d, a.size, d2, b.size)
        ctx.device_free(d); ctx.device_free(d2)
    else:
        bins, nt, nd = job.hist_job(a.ctypes.data, a.size, device=False)
        mx, _, _ = job.gcp_job(a.ctypes.data, a.size, 0.5, 300, device=False)
        cmp_ = job.comp_job(a.ctypes.data, a.size, b.ctypes.data, b.size, device=False)
    res[mode] = (bins, mx, cmp_, nt, nd)
st = ctx.dist_stats()
# every rank holds the complete result
for mode, (bins, mx, cmp_, nt, nd) in res.items():
    np.savez(%(out)r + ".%%s.%%d.npz" %% (mode, rank), bins=bins, mx=mx, main=cmp_["main"], spectra=cmp_["spectra"],
             counters=np.array([cmp_["counters"][f] for f in B.CC_FIELDS], dtype=np.uint64), tot=np.array([nt, nd], dtype=np.uint64),
             rounds=np.array([st["rounds"]]))
ctx.close(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("even", [1, 0], ids=["even_shards", "ragged_and_empty_shards"])
def test_two_rank_job_matches_oracle(tmp_path, even):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import oracle as O
    from kat_b200 import binding as B
    out = str(tmp_path / "r")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": out, "even": even})
    env = dict(os.environ, KATGPU_DIST_ROUND_MB="2")  # 9 MB shards: several route rounds, both receive areas reused
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)], timeout=900, env=env)
    a = B.synth_reads_host(300_000, 7, 8, 0, 60000, 150)
    if even:
        b = np.concatenate([B.synth_reads_host(300_000, 7, 9, 0, 15000, 150, 0.02, 0.001),
                            B.synth_reads_host(300_000, 7, 9, 15000, 15000, 150, 0.02, 0.001)])
    else:
        b = np.concatenate([B.synth_reads_host(300_000, 7, 9, 0, (60000 - 7) // 2, 150, 0.02, 0.001),
                            B.synth_reads_host(300_000, 7, 9, (60000 - 7) // 2, 3, 150, 0.02, 0.001)])
    oa, ob = O.Table.from_chunks([a.tobytes()], 27, True), O.Table.from_chunks([b.tobytes()], 27, True)
    oc = O.comp(oa, ob)
    for mode in ("device", "host"):
        for rank in (0, 1):
            r = np.load(out + ".%s.%d.npz" % (mode, rank))
            assert int(r["tot"][0]) == oa.total and int(r["tot"][1]) == oa.distinct
            assert np.array_equal(r["bins"], oa.hist()[1])
            assert np.array_equal(r["mx"], oa.gcp(0.5, 300))
            assert np.array_equal(r["main"], oc["main"]) and np.array_equal(r["spectra"], oc["spectra"])
            assert [int(x) for x in r["counters"]] == [oc["counters"][f] for f in B.CC_FIELDS]
            assert int(r["rounds"][0]) >= 10


@pytest.mark.parametrize("case", [c for c in G.cases() if c["name"] in
                                  ("hist_two_files_k27", "gcp_k17", "comp_k15_glob_default_bins", "comp_k17_three_inputs", "hist_k32_gz", "hist_polyA")],
                         ids=lambda c: c["name"])
def test_cli_two_gpus_byte_identical(case, tmp_path):
    """`kat --gpus 2 hist|gcp|comp`: one process, one host thread and one context per GPU, NCCL
    all-reduce of the result; the files equal the reference's byte for byte."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    shutil.copytree(os.path.join(G.GOLDEN, "inputs"), tmp_path / "inputs")
    out = os.path.join("expected", case["name"])
    os.makedirs(tmp_path / out)
    args = [a.replace("OUT", out) for a in case["args"]]
    env = dict(os.environ, KATGPU_DIST_ROUND_MB="1")
    r = subprocess.run([KAT, args[0], "--gpus", "2"] + args[1:], cwd=tmp_path, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    for f in case["files"]:
        got = open(tmp_path / out / f, "rb").read()
        assert got == G.expected_bytes(case["name"], f), f"{case['name']}/{f} differs from the reference's file"
