"""BASELINE configs 4 and 5 at size, through the drop-in, checked against the oracle (VERDICT r1 next #1d).

C4  MiniLM-shaped text table (5M x 384) + CLIP-shaped image table (1M x 512), both loaded with `B200Store.load_arrow`
    (columnar host copy, scatter loader), searched with `fused_search_batch` (text top-50 + image top-12 + device
    z-score fusion + FINAL_N + CONFIDENCE_TAU): every request's scans are compared with the oracle's threaded flat search
    on the original fp32 rows (tolerance-aware), and the device fusion must equal the oracle fusion of the device's own
    scan output bit for bit.
C5  1000 tenants of ragged size in ONE varlen launch: every query's result equals the oracle on that tenant's rows.
"""
import importlib
import os
import sys
import time

import numpy as np
import pytest
import torch

from oracle import flat_search as ofs
from oracle import fusion as ofu
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "multimodal-rag-for-image-text-search_b200"


def _ids(prefix, n):
    import pyarrow as pa
    import pyarrow.compute as pc
    return pc.binary_join_element_wise(pa.array([prefix] * 1, pa.string()).take(pa.array(np.zeros(n, np.int32))),
                                       pa.array(np.arange(n)).cast(pa.string()), "")


def _const(value, n):
    import pyarrow as pa
    return pa.array([value], pa.string()).take(pa.array(np.zeros(n, np.int32)))


def test_c4_fused_text_plus_image_at_size():
    import bench
    import pyarrow as pa

    pkg = importlib.import_module(PKG)
    if torch.cuda.mem_get_info()[0] < 30 * 2 ** 30 or bench.host_mem_available_gb() < 40:
        pytest.skip("needs ~12 GB of HBM and ~25 GB of host memory")
    n_t, n_i, cores = 5_000_000, 1_000_000, bench.host_cores()
    temb, iemb = bench.host_table(n_t, 384, cores), bench.host_table(n_i, 512, cores)
    # two tenants interleaved in insertion order: the loader must scatter each block into tenant-sorted HBM rows
    ut = np.where(np.arange(n_t) % 5 == 0, "bob", "alice")
    ui = np.where(np.arange(n_i) % 4 == 0, "bob", "alice")
    store = pkg.B200Store()
    t0 = time.perf_counter()
    store.load_arrow("text_collection", pkg.make_arrow_table(_ids("t", n_t), pa.array(ut), _const("d", n_t), _const("text", n_t),
                                                             temb, _const("{}", n_t)))
    store.load_arrow("image_collection", pkg.make_arrow_table(_ids("i", n_i), pa.array(ui), _const("d", n_i), _const("image", n_i),
                                                              iemb, _const("{}", n_i)))
    users = ["alice", "bob", "alice", "nobody", "bob"]
    qt, qi = util.queries(len(users), 384, seed=61), util.queries(len(users), 512, seed=62)
    got = store.fused_search_batch(users, qt, qi, 50, 12, 4, 0.25)
    load_s = time.perf_counter() - t0
    assert store._text_table.rebuilds == 1 and len(store._text_table) == n_t and store._text_table.last_load_gbs > 1.0
    print(f"C4 load+first search {load_s:.1f} s; loader {store._text_table.last_load_gbs:.1f} GB/s (fp32 source bytes)")
    tsel = {"alice": np.nonzero(ut == "alice")[0], "bob": np.nonzero(ut == "bob")[0]}
    isel = {"alice": np.nonzero(ui == "alice")[0], "bob": np.nonzero(ui == "bob")[0]}
    for j, u in enumerate(users):
        items, low = got[j]
        if u == "nobody":
            assert items == [] and low is True
            continue
        # the two scans, through the public per-request calls, against the oracle on that tenant's fp32 rows
        for coll, emb, sel, q, k, call in ((store._text_table, temb, tsel[u], qt[j], 50, store.search_text),
                                           (store._image_table, iemb, isel[u], qi[j], 12, store.search_image)):
            hits = call(u, q.tolist(), k)
            full = util.oracle_scores(emb[sel], q)
            rows = np.array([int(h["chunk_id"][1:]) for h in hits])
            assert np.isin(rows, sel).all(), "hit outside the tenant"
            local = np.searchsorted(sel, rows)
            util.check_topk(np.array([h["score"] for h in hits], np.float32), local, full, k, util.TOL_BF16, what=f"C4 {coll.name} {u}")
        # device fusion == oracle fusion of the device's own scan output, bit for bit
        text = [{"chunk_id": h["chunk_id"], "score": h["score"]} for h in store.search_text(u, qt[j].tolist(), 50)]
        image = [{"chunk_id": h["chunk_id"], "score": h["score"]} for h in store.search_image(u, qi[j].tolist(), 12)]
        want = ofu.fuse_results(text, image, 4)
        assert [it["chunk_id"] for it in items] == [w["chunk_id"] for w in want]
        assert [it["combined_score"] for it in items] == [w["combined_score"] for w in want]
        assert low is ofu.confidence_low(want, 0.25)


def test_c5_thousand_tenants_one_varlen_launch():
    import bench

    pkg = importlib.import_module(PKG)
    if torch.cuda.mem_get_info()[0] < 30 * 2 ** 30:
        pytest.skip("needs ~10 GB of free HBM")
    rng = np.random.default_rng(7)
    sizes = np.exp(rng.uniform(np.log(100), np.log(30_000), size=1000)).astype(np.int64)   # ragged, log-uniform
    sizes[17], sizes[400] = 1, 0                                                           # a one-row and an empty tenant
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    total = int(seg[-1])
    base = bench.build_shard(pkg, 0, total, 512, "bf16", torch.device("cuda:0"))
    ix = pkg.ResidentIndex(base.rows, seg_offsets=seg)
    b = 1000
    tenants = rng.permutation(1000).astype(np.int32)                                       # every tenant once
    tenants[:8] = tenants[8]                                                               # ... and one tenant 9 times
    q = util.queries(b, 512, seed=71)
    s, r = ix.search(torch.from_numpy(q).cuda(), 10, tenants)
    assert pkg._native.lib().mmr_last_kernel() == 3
    s, r = s.cpu().numpy(), r.cpu().numpy()
    for j in range(b):
        t = int(tenants[j])
        lo, hi = int(seg[t]), int(seg[t + 1])
        rows = base.rows[lo:hi].float().cpu().numpy()
        if hi == lo:
            assert (r[j] == -1).all() and np.isneginf(s[j]).all()
            continue
        # strict mode: the oracle on exactly the stored bf16 values must give identical ids
        util.check_topk(s[j], r[j], util.oracle_scores(rows, q[j]), 10, util.TOL_STRICT * 5, lo=lo, hi=hi, what=f"C5 tenant {t}")
    ix.close()
    base.close()
