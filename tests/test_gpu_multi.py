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
            pkg._native.set_option("MMR_PDL", "1" if mode == "fused-pipelined" else "0")
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
        # host-buffer form of the fused exchange (query in the kernel parameters, merged result in a mapped mailbox)
        pkg._native.set_option("MMR_PDL", "0")
        sh = pkg.ShardedIndex(local, exchange="fused")
        for b, k in ((1, 10), (2, 12), (5, 10)):
            qs = np.concatenate([util.queries(b - 1, 512, seed=50 + b), rows[11:12]]) if b > 1 else rows[11:12].copy()
            for rep in range(3):
                hs, hr = sh.search_host(qs, k)
            s1, r1 = single.search(torch.from_numpy(qs).cuda(), k)
            assert (hr == r1.cpu().numpy()).all() and (hs == s1.cpu().numpy()).all(), f"search_host b{b} rank{rank}"
        # tenants: segments that straddle shard boundaries keep their id on every shard that holds a slice of them
        sharded_mod = importlib.import_module(PKG + ".sharded")
        seg = np.array([0, 1000, 1000, 140_000, 260_000, n], dtype=np.int64)
        local_seg = pkg.ResidentIndex(full_dev[lo:hi].contiguous(), seg_offsets=sharded_mod.split_segments(seg, lo, hi), row_base=lo)
        single_seg = pkg.ResidentIndex(full_dev, seg_offsets=seg)
        pkg._native.set_option("MMR_PDL", "0")
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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_multi_index_in_one_process_equals_single_gpu():
    """One process, G GPUs (mmr_multi_*): per-device launcher threads, scan kernels push into the collector's buffer over
    NVLink peer mappings, merged result in a mapped mailbox.  Bit-identical to the single-GPU scan for every shape."""
    pkg = importlib.import_module(PKG)
    G = min(torch.cuda.device_count(), 8)
    n = 400_003
    rows = util.unit_rows(n, 512, seed=91)
    rows[n - 3] = rows[5]                                     # a tie between the first and the last shard
    bounds = pkg.shard_bounds(n, G)
    single = pkg.ResidentIndex.from_f32(rows, dtype="bf16", device="cuda:0")
    shards = [pkg.ResidentIndex(single.rows[bounds[g]:bounds[g + 1]].to(f"cuda:{g}").contiguous(), row_base=bounds[g])
              for g in range(G)]
    multi = pkg.MultiIndex(shards)
    q = np.concatenate([rows[5:6], util.queries(8, 512, seed=92)])
    for b, k in ((1, 10), (2, 12), (1, 50), (5, 10), (9, 10)):
        for rep in range(4):                                  # sequence numbers / slot parity roll over
            hs, hr = multi.search_host(q[:b], k)
        ds, dr = single.search(torch.from_numpy(q[:b]).cuda(), k)
        assert (hr == dr.cpu().numpy()).all() and (hs == ds.cpu().numpy()).all(), (b, k)
        assert hr[0, 0] == 5 and hr[0, 1] == n - 3
    # ragged ranges per query: ranges that straddle shard boundaries, empty ranges, several ranges
    ranges = [[(0, n)], [(bounds[1] - 1000, bounds[1] + 1000)], [], [(10, 20), (bounds[-2] - 5, n)], [(7, 8)]]
    hs, hr = multi.search_host(q[:5], 10, ranges)
    ds, dr = single.search_ranges(torch.from_numpy(q[:5]).cuda(), 10, ranges)
    assert (hr == dr.cpu().numpy()).all() and (hs == ds.cpu().numpy()).all()
    multi.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_store_behind_the_drop_in_equals_single_gpu_store():
    """B200Store(devices=[...]): the same search_* calls, row-range shards on every GPU, identical dicts to one GPU --
    also after upserts (delta segments on the last shard, tombstones on whichever shard holds the replaced row)."""
    pkg = importlib.import_module(PKG)
    G = min(torch.cuda.device_count(), 8)
    rng = np.random.default_rng(93)
    n = 60_000
    emb = util.unit_rows(n, 384, seed=94)
    users = [f"u{i % 5}" for i in range(n)]
    table = pkg.make_arrow_table([f"t{i}" for i in range(n)], users, ["d"] * n, ["text"] * n, emb, ["{}"] * n)
    one, many = pkg.B200Store(), pkg.B200Store(devices=list(range(G)))
    for st in (one, many):
        st.load_arrow("text_collection", table)

    def check(tag):
        qs = rng.standard_normal((7, 384)).astype(np.float32)
        us = ["u0", "u3", "u3", "nobody", "u1", "u3", "u4"]
        for u, q in zip(us, qs):
            assert many.search_text(u, q.tolist(), 10) == one.search_text(u, q.tolist(), 10), (tag, u)
        assert many.search_text_batch(us, qs, 50) == one.search_text_batch(us, qs, 50), tag

    check("loaded")
    assert many._text_table._multi is not None and len(many._text_table._shards) == G
    for step in range(3):
        new = [pkg.VectorRow(f"n{step}_{i}", f"u{i % 5}", "d", "text", rng.standard_normal(384).tolist(), {"s": step}) for i in range(30)]
        over = [pkg.VectorRow(f"t{int(i)}", users[int(i)], "d", "text", rng.standard_normal(384).tolist(), {"o": step})
                for i in rng.choice(n, size=20, replace=False)]
        for st in (one, many):
            st.upsert_text_vectors(new + over)
        check(f"upsert{step}")
    assert many._text_table.rebuilds == 1 and many._text_table.appends == 3
