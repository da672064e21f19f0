"""Seeded fuzz of the whole search surface against the oracle: random table sizes, dims, storage types, batch sizes,
k, tenant layouts and query->tenant assignments (uniform launches, tensor-core launches, ragged launches, explicit
row ranges).  Every result goes through the tolerance-aware checker."""
import importlib

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu
PKG = "multimodal-rag-for-image-text-search_b200"
TOL = {"bf16": util.TOL_BF16, "f16": util.TOL_BF16, "f32": util.TOL_F32}


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_against_oracle(seed):
    mmr = importlib.import_module(PKG)
    rng = np.random.default_rng(1000 + seed)
    dim = int(rng.choice([384, 512]))
    dtype = str(rng.choice(["bf16", "bf16", "f16", "f32"]))
    n = int(rng.integers(1, 6000)) if seed % 3 else int(rng.integers(20_000, 90_000))
    n_seg = int(rng.integers(1, 9))
    cuts = np.sort(rng.integers(0, n + 1, size=n_seg - 1)) if n_seg > 1 else np.array([], dtype=np.int64)
    seg = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    rows = util.unit_rows(n, dim, seed=seed, cone=float(rng.choice([0.0, 0.3])))
    if n > 40:                                   # a few exact duplicates -> ties
        for _ in range(3):
            a, b = rng.integers(0, n, size=2)
            rows[a] = rows[b]
    ix = mmr.ResidentIndex.from_f32(rows, seg_offsets=seg, dtype=dtype)
    stored = ix.rows.float().cpu().numpy()
    for trial in range(4):
        b = int(rng.integers(1, 13))
        k = int(rng.choice([1, 3, 10, 12, 50, 64]))
        q = util.queries(b, dim, seed=seed * 10 + trial) * np.float32(rng.uniform(0.1, 5.0))
        mode = trial % 4
        if mode == 0:
            segs = None                                                    # whole table, uniform
            ranges = [[(0, n)]] * b
        elif mode == 1:
            t = int(rng.integers(0, n_seg))
            segs = [t] * b                                                 # one tenant, uniform (K1 or K2 by batch size)
            ranges = [[(int(seg[t]), int(seg[t + 1]))]] * b
        elif mode == 2:
            segs = rng.integers(-1, n_seg, size=b).tolist()                # mixed tenants -> ragged launch
            ranges = [[(0, n)] if t < 0 else [(int(seg[t]), int(seg[t + 1]))] for t in segs]
        else:
            segs = "ranges"                                                # explicit multi-range queries
            ranges = []
            for _ in range(b):
                nr = int(rng.integers(0, 4))                   # 0..3 disjoint ranges (overlaps are rejected by the ABI)
                cuts = np.sort(rng.integers(0, n + 1, size=2 * nr))
                rr = [(int(cuts[2 * i]), int(cuts[2 * i + 1])) for i in range(nr)]
                rng.shuffle(rr)                                 # any order
                ranges.append([tuple(x) for x in rr])
        qd = torch.from_numpy(q).cuda()
        if segs == "ranges":
            s, r = ix.search_ranges(qd, k, ranges)
        else:
            s, r = ix.search(qd, k, segs)
        s, r = s.cpu().numpy(), r.cpu().numpy()
        family = mmr._native.lib().mmr_last_kernel()
        for j in range(b):
            # oracle over the union of the query's ranges (overlapping ranges may list a row twice: de-duplicate)
            rowset = np.unique(np.concatenate([np.arange(lo, hi) for lo, hi in ranges[j]] + [np.array([], dtype=np.int64)])).astype(np.int64)
            what = f"seed{seed} trial{trial} dtype={dtype} d={dim} n={n} b={b} k={k} mode={mode} fam={family} q{j}"
            if segs == "ranges" and any(hi > lo for lo, hi in ranges[j]) and len(ranges[j]) > 1:
                overl = sum(hi - lo for lo, hi in ranges[j]) != len(rowset)
                if overl:
                    continue                     # overlapping explicit ranges are a caller error (rows counted twice)
            full = util.oracle_scores(stored[rowset], q[j]) if len(rowset) else np.zeros(0, np.float32)
            tol = TOL[dtype] if family == 2 else util.TOL_STRICT * 5   # K1 on the stored values: fp32 summation order only
            if len(rowset) == 0:
                assert (r[j] == -1).all(), what
                continue
            # map row ids -> positions inside rowset for the checker
            pos = np.searchsorted(rowset, np.where(r[j] >= 0, r[j], rowset[0]))
            assert ((r[j] < 0) | (rowset[np.clip(pos, 0, len(rowset) - 1)] == r[j])).all(), what + ": row outside its ranges"
            rj = np.where(r[j] >= 0, pos, -1)
            util.check_topk(s[j], rj, full, k, tol, what=what)
    ix.close()
