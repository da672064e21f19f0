"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo processes run the shard bounds,
the packed wire buffer and the all-gather exactly as the CUDA path does; the oracle stands in for the
per-shard scan (test infrastructure) and the gathered result must equal the single-scan oracle result."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flat_search as ofs
from tests import util

PKG = "multimodal-rag-for-image-text-search_b200"


def merge_gathered_numpy(gathered):
    """Host statement of the K4 merge, used only here to check the exchange logic on CPU (the product merges with
    the CUDA kernel mmr_merge_topk_strided)."""
    scores, rows = gathered.views()
    w = gathered.wire
    s = scores.cpu().numpy().reshape(gathered.world, w.b, w.k)
    r = rows.cpu().numpy().reshape(gathered.world, w.b, w.k)
    out_s = np.full((w.b, w.k), -np.inf, np.float32)
    out_r = np.full((w.b, w.k), -1, np.int64)
    for q in range(w.b):
        cs, cr = s[:, q].reshape(-1), r[:, q].reshape(-1)
        keep = cr >= 0
        cs, cr = cs[keep], cr[keep]
        order = np.lexsort((cr, -cs.astype(np.float64)))[: w.k]
        out_s[q, : len(order)], out_r[q, : len(order)] = cs[order], cr[order]
    return out_s, out_r


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, dim, b, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = importlib.import_module(PKG + ".sharded")
        rows = util.unit_rows(n_rows, dim, seed=5)
        rows[n_rows - 1] = rows[3]                 # a tie across the first and last shard
        qs = np.concatenate([util.queries(b - 1, dim), rows[3:4]])
        bounds = sh.shard_bounds(n_rows, world, align=8)
        lo, hi = bounds[rank], bounds[rank + 1]
        wire = sh.Wire(b, k, "cpu")
        gathered = sh.GatheredWire(wire, world)
        # the per-shard scan (CUDA in production) restated by the oracle: every rank derives the same full
        # distance matrix (BLAS results depend on the operand shape, so slices of ONE product keep ties exact)
        # and keeps the best k of its own row range, as global ids (row_base = lo)
        qn = np.stack([np.asarray(ofs.normalize(v), np.float32) for v in qs])
        dist_all = (np.float32(1) - (rows @ qn.T).astype(np.float32)).T
        wire.scores.fill_(float("-inf"))
        wire.rows.fill_(-1)
        for q in range(b):
            ids = ofs.topk_smallest(dist_all[q, lo:hi], k)
            wire.scores[q, :len(ids)] = torch.from_numpy((np.float32(1) - dist_all[q, lo:hi][ids]).astype(np.float32))
            wire.rows[q, :len(ids)] = torch.from_numpy(ids + lo)
        sh.gather_wire(wire, gathered)
        s, r = merge_gathered_numpy(gathered)
        if rank == 0:
            ret["bounds"] = bounds
            ret["scores"], ret["rows"] = s, r
            want_i = np.stack([ofs.topk_smallest(dist_all[q], k) for q in range(b)])
            ret["want_rows"] = want_i
            ret["want_scores"] = np.stack([(np.float32(1) - dist_all[q][want_i[q]]).astype(np.float32) for q in range(b)])
        # every rank holds the same merged answer
        t = torch.from_numpy(r.copy())
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows", [(2, 4001), (3, 1000), (2, 9)])
def test_shard_gather_merge_equals_single_scan(world, n_rows):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_rows, 384, 4, 10, ret), nprocs=world, join=True)
    assert ret["bounds"][0] == 0 and ret["bounds"][-1] == n_rows and len(ret["bounds"]) == world + 1
    kk = ret["want_rows"].shape[1]
    assert (ret["rows"][:, :kk] == ret["want_rows"]).all()
    assert np.array_equal(ret["scores"][:, :kk], ret["want_scores"])
    assert (ret["rows"][:, kk:] == -1).all()
    # the duplicated row: both copies tie, lower ordinal first
    assert ret["rows"][3, 0] == 3 and ret["rows"][3, 1] == n_rows - 1


def test_bounds_and_segment_split():
    sh = importlib.import_module(PKG + ".sharded")
    for n, w in ((10_000_000, 8), (10, 3), (0, 2), (7, 8)):
        b = sh.shard_bounds(n, w, align=8)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
    b = sh.shard_bounds(10_000_000, 8)
    assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1
    seg = [0, 5, 5, 40, 100]
    assert sh.split_segments(seg, 0, 30).tolist() == [0, 5, 5, 30, 30]
    assert sh.split_segments(seg, 30, 100).tolist() == [0, 0, 0, 10, 70]


def test_wire_layout():
    sh = importlib.import_module(PKG + ".sharded")
    w = sh.Wire(3, 5, "cpu")                      # 15 scores = 60 B -> padded to 64 so the rows are 8-aligned
    assert w.score_bytes == 64 and w.nbytes == 64 + 120
    w.scores.copy_(torch.arange(15, dtype=torch.float32).view(3, 5))
    w.rows.copy_(torch.arange(15, dtype=torch.int64).view(3, 5) + 100)
    g = sh.GatheredWire(w, 2)
    g.buf[0].copy_(w.buf)
    g.buf[1].copy_(w.buf)
    s, r = g.views()
    assert s.shape == (2, 15) and r.shape == (2, 15)
    assert s.stride(0) * 4 == w.nbytes and r.stride(0) * 8 == w.nbytes
    assert r[1, 14].item() == 114 and s[1, 14].item() == 14.0
