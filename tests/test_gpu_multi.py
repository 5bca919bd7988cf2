"""N>1 on real GPUs: 2 ranks under torchrun (NCCL), partitioned counting must
reproduce the single-table oracle result.  Skipped on a 1-GPU box."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import kat_b200
from kat_b200 import binding as B
from kat_b200.multigpu import DistributedCounter
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = kat_b200.Context(lr)
n_reads, L, k = 30000, 150, 27
per = n_reads // world
host = B.synth_reads_host(300_000, 7, 8, rank * per, per, L)
d = ctx.device_alloc(host.size)
ctx.h2d(d, host)
dc = DistributedCounter(ctx, dist, k, True, 1 << 20, batch_bytes=(L + 1) * 16 * 300)
bins, nt, nd = dc.hist_job(d, host.size, record_len=L + 1)
mx, nt2, nd2 = dc.gcp_job(d, host.size, record_len=L + 1)
tot = torch.tensor([nt, nd], device="cuda"); dist.all_reduce(tot)
if rank == 0:
    np.savez(%(out)r, bins=bins, mx=mx, tot=tot.cpu().numpy())
ctx.close(); dist.destroy_process_group()
'''


def test_two_rank_partitioned_hist_and_gcp(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import oracle as O
    from kat_b200 import binding as B
    out = str(tmp_path / "r0.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": out})
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)], timeout=600)
    r = np.load(out)
    full = B.synth_reads_host(300_000, 7, 8, 0, 30000, 150)
    ot = O.Table.from_chunks([full.tobytes()], 27, True)
    assert int(r["tot"][0]) == ot.total and int(r["tot"][1]) == ot.distinct
    assert np.array_equal(r["bins"], ot.hist()[1])
    assert np.array_equal(r["mx"], ot.gcp())
