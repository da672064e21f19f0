"""Real multi-GPU path (needs >= 2 GPUs: `gpurun --gpus 2`): row-range shards, NCCL all-gather exchange and the fused
peer-memory exchange must both reproduce the single-GPU scan bit for bit."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util

pytestmark = pytest.mark.gpu
PKG = "multimodal-rag-for-image-text-search_b200"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pkg = importlib.import_module(PKG)
        n = 300_000
        rows = util.unit_rows(n, 512, seed=77)
        rows[n - 5] = rows[11]                                # a tie across the first and the last shard
        full_dev = torch.from_numpy(rows).cuda().to(torch.bfloat16)
        bounds = pkg.shard_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        local = pkg.ResidentIndex(full_dev[lo:hi].contiguous(), row_base=lo)
        single = pkg.ResidentIndex(full_dev)
        results = {}
        for mode in ("nccl", "fused", "fused-pipelined"):
            os.environ["MMR_PDL"] = "1" if mode == "fused-pipelined" else "0"
            sh = pkg.ShardedIndex(local, exchange="fused" if mode.startswith("fused") else "nccl")
            for b, k in ((1, 10), (2, 12), (5, 10), (130, 50)):
                qs = np.concatenate([util.queries(b - 1, 512, seed=b), rows[11:12]]) if b > 1 else rows[11:12].copy()
                qd = torch.from_numpy(qs).cuda()
                for rep in range(3 if mode != "fused-pipelined" else 40):  # sequence numbers / slot parity roll over
                    s, r = sh.search(qd, k)
                torch.cuda.synchronize()
                s1, r1 = single.search(qd, k)
                assert torch.equal(r, r1), f"{mode} b{b} k{k} rank{rank}: ids differ from the single-GPU scan"
                assert torch.equal(s, s1), f"{mode} b{b} k{k} rank{rank}: scores differ"
                assert r[-1, 0].item() == 11 and r[-1, 1].item() == n - 5
                results[(mode, b, k)] = r.cpu()
        for (mode, b, k), r in results.items():
            if mode != "nccl":
                assert torch.equal(r, results[("nccl", b, k)])
        # tenants: segments that straddle shard boundaries keep their id on every shard that holds a slice of them
        sharded_mod = importlib.import_module(PKG + ".sharded")
        seg = np.array([0, 1000, 1000, 140_000, 260_000, n], dtype=np.int64)
        local_seg = pkg.ResidentIndex(full_dev[lo:hi].contiguous(), seg_offsets=sharded_mod.split_segments(seg, lo, hi), row_base=lo)
        single_seg = pkg.ResidentIndex(full_dev, seg_offsets=seg)
        os.environ["MMR_PDL"] = "0"
        for mode in ("nccl", "fused"):
            sh = pkg.ShardedIndex(local_seg, exchange=mode)
            for tenants in ([2], [3, 3, 3, 3, 3], [0, 2, 3, 4, 1, -1]):
                b = len(tenants)
                qd = torch.from_numpy(util.queries(b, 512, seed=900 + b)).cuda()
                s, r = sh.search(qd, 10, tenants)
                torch.cuda.synchronize()
                s1, r1 = single_seg.search(qd, 10, tenants)
                assert torch.equal(r, r1) and torch.equal(s, s1), f"{mode} tenants {tenants} rank{rank}"
                for j, t in enumerate(tenants):
                    if t >= 0:
                        ok = (r[j] < 0) | ((r[j] >= int(seg[t])) & (r[j] < int(seg[t + 1])))
                        assert bool(ok.all())
        if rank == 0:
            ret["ok"] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_exchange_matches_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret.get("ok")
