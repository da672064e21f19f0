"""Parity of the CUDA scan (through the C ABI) against the CPU oracle.  Run with -m gpu on a B200."""
import importlib

import numpy as np
import pytest
import torch

from oracle import flat_search as ofs
from tests import util

pytestmark = pytest.mark.gpu
PKG = "multimodal-rag-for-image-text-search_b200"


@pytest.fixture(scope="module")
def mmr():
    pkg = importlib.import_module(PKG)
    pkg._native.lib()  # the CUDA extension must be the thing that runs
    assert torch.cuda.is_available()
    return pkg


def _search_both_ways(ix, q_np, k, segments=None):
    dev = ix.device
    s_dev, r_dev = ix.search(torch.from_numpy(q_np).to(dev), k, segments)
    torch.cuda.synchronize()
    s_host, r_host = ix.search_host(q_np, k, segments)
    assert (s_dev.cpu().numpy() == s_host).all() or np.array_equal(s_dev.cpu().numpy(), s_host, equal_nan=True)
    assert (r_dev.cpu().numpy() == r_host).all()
    return s_host, r_host


@pytest.mark.parametrize("dtype,tol", [("bf16", util.TOL_BF16), ("f16", util.TOL_BF16), ("f32", util.TOL_F32)])
@pytest.mark.parametrize("dim", [384, 512])
def test_small_table_all_dtypes(mmr, dtype, tol, dim):
    """C1-sized table (10k rows): every storage dtype, k in {1, 10, 12, 50, 64}, B in {1..5, 9}."""
    rows = util.unit_rows(10_000, dim, seed=dim)
    ix = mmr.ResidentIndex.from_f32(rows, dtype=dtype)
    qs = util.queries(9, dim) * np.float32(2.5)  # search re-normalises (lancedb_store.py:104)
    full = [util.oracle_scores(rows, q) for q in qs]
    for k in (1, 10, 12, 50, 64):
        for b in (1, 2, 3, 4, 5, 9):
            s, r = _search_both_ways(ix, qs[:b], k)
            for j in range(b):
                util.check_topk(s[j], r[j], full[j], k, tol, what=f"{dtype} d{dim} k{k} b{b} q{j}")
    ix.close()


def test_strict_ids_on_stored_values(mmr):
    """Oracle evaluated on exactly the bf16 values the index holds: ids identical (up to fp32 summation-order
    ties), scores within 2e-6 -- isolates kernel bugs from quantisation."""
    rows = util.unit_rows(200_000, 512, seed=11)
    ix = mmr.ResidentIndex.from_f32(rows, dtype="bf16")
    stored = ix.rows.to(torch.float32).cpu().numpy()
    assert (stored == ofs.bf16_round(rows)).all(), "loader must round to nearest even"
    qs = util.queries(2, 512)            # B <= 2 -> K1 (fp32 queries); K2's strict check lives in test_gpu_umma.py
    s, r = _search_both_ways(ix, qs, 10)
    assert mmr._native.lib().mmr_last_kernel() == 1
    for j in range(2):
        full = util.oracle_scores(stored, qs[j])
        util.check_topk(s[j], r[j], full, 10, util.TOL_STRICT, what=f"strict q{j}")
        d, ids = ofs.flat_search(stored, qs[j], 10)
        gap = np.diff(-(1.0 - d.astype(np.float64))).min()
        if gap > 4 * util.TOL_STRICT:
            assert r[j].tolist() == ids.tolist()
    ix.close()


def test_config2_1m_x_512(mmr):
    """BASELINE config 2: CLIP-shaped 1M x 512 bf16 index, top-10, small batches (K1 path)."""
    rows = util.unit_rows(1_000_000, 512, seed=21, cone=0.3)   # index B: CLIP-like cone, small score gaps
    ix = mmr.ResidentIndex.from_f32(rows, dtype="bf16")
    qs = util.queries(4, 512, cone=0.3)
    for b in (1, 4):
        s, r = _search_both_ways(ix, qs[:b], 10)
        for j in range(b):
            util.check_topk(s[j], r[j], util.oracle_scores(rows, qs[j]), 10, util.TOL_BF16, what=f"1M b{b} q{j}")
    ix.close()


def test_tie_rule_exact_duplicates(mmr):
    """0.1 % exact duplicate rows: equal scores must come out ordered by row id, and the k boundary must
    keep the smaller row ids."""
    rows = util.unit_rows(50_000, 512, seed=31)
    q = util.queries(1, 512)[0]
    best = int(np.argmax(rows @ q))
    dup_at = [17, 4242, 4243, 30_001, 49_999]
    for p in dup_at:
        rows[p] = rows[best]
    ix = mmr.ResidentIndex.from_f32(rows, dtype="bf16")
    expect = sorted(set(dup_at + [best]))
    for k in (3, 6, 10):
        s, r = _search_both_ways(ix, q[None], k)
        m = min(k, len(expect))
        assert r[0][:m].tolist() == expect[:m]
        assert (s[0][:m] == s[0][0]).all()
    ix.close()


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 63, 1000, 1185, 9473])
def test_ragged_sizes_and_fewer_rows_than_k(mmr, n):
    rows = util.unit_rows(max(n, 1), 384, seed=41)[:n]
    dev_rows = torch.from_numpy(ofs.bf16_round(rows)).to("cuda", torch.bfloat16).reshape(n, 384)
    ix = mmr.ResidentIndex(dev_rows)
    qs = util.queries(2, 384)
    s, r = _search_both_ways(ix, qs, 10)
    for j in range(2):
        if n == 0:
            assert (r[j] == -1).all() and np.isneginf(s[j]).all()
        else:
            util.check_topk(s[j], r[j], util.oracle_scores(ofs.bf16_round(rows), qs[j]), 10, util.TOL_STRICT, what=f"n{n}")
    ix.close()


def test_tenant_segments_prefilter(mmr):
    """Tenant = contiguous row segment; results never leave it; unknown/empty tenants give no hits; a batch
    whose queries hit different tenants goes through the varlen launch and matches per-tenant scans."""
    sizes = [3, 0, 1500, 8, 20_000, 1, 977]
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    rows = util.unit_rows(int(seg[-1]), 512, seed=51)
    ix = mmr.ResidentIndex.from_f32(rows, seg_offsets=seg, dtype="bf16")
    qs = util.queries(len(sizes), 512)
    # one tenant per call (uniform launch)
    singles = []
    for t, n in enumerate(sizes):
        s, r = _search_both_ways(ix, qs[t:t + 1], 12, [t])
        singles.append((s[0], r[0]))
        if n == 0:
            assert (r[0] == -1).all()
        else:
            full = util.oracle_scores(rows[seg[t]:seg[t + 1]], qs[t])
            util.check_topk(s[0], r[0], full, 12, util.TOL_BF16, lo=int(seg[t]), hi=int(seg[t + 1]), what=f"tenant{t}")
    # all tenants in one ragged launch
    s, r = _search_both_ways(ix, qs, 12, list(range(len(sizes))))
    assert mmr._native.lib().mmr_last_kernel() == 3
    for t in range(len(sizes)):
        assert r[t].tolist() == singles[t][1].tolist()
        assert np.array_equal(s[t], singles[t][0])
    # same tenant for every query -> uniform launch again; -1 = whole table
    s, r = _search_both_ways(ix, qs[:3], 5, [4, 4, 4])
    assert mmr._native.lib().mmr_last_kernel() in (1, 2)
    s_all, r_all = _search_both_ways(ix, qs[:1], 5, [-1])
    util.check_topk(s_all[0], r_all[0], util.oracle_scores(rows, qs[0]), 5, util.TOL_BF16, what="whole table")
    ix.close()


def test_varlen_many_tenants(mmr):
    """Config-5 shape at test size: 200 ragged tenants, one query each plus repeats, one launch."""
    rng = np.random.default_rng(7)
    sizes = np.exp(rng.uniform(np.log(10), np.log(20_000), size=200)).astype(np.int64)
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    rows = util.unit_rows(int(seg[-1]), 512, seed=52)
    ix = mmr.ResidentIndex.from_f32(rows, seg_offsets=seg, dtype="bf16")
    tenants = np.concatenate([np.arange(200), rng.integers(0, 200, size=56)]).astype(np.int32)
    qs = util.queries(len(tenants), 512)
    s, r = _search_both_ways(ix, qs, 10, tenants)
    for j, t in enumerate(tenants):
        full = util.oracle_scores(rows[seg[t]:seg[t + 1]], qs[j])
        util.check_topk(s[j], r[j], full, 10, util.TOL_BF16, lo=int(seg[t]), hi=int(seg[t + 1]), what=f"varlen q{j} t{t}")
    ix.close()


def test_row_base_and_shard_merge_equals_single_scan(mmr):
    """Row-range shards + K4 merge give bit-identical results to one scan, for G in {2, 4, 8}."""
    rows = util.unit_rows(120_000, 512, seed=61)
    rows[100_000] = rows[5]          # a cross-shard tie
    ix = mmr.ResidentIndex.from_f32(rows, dtype="bf16")
    qs = np.concatenate([util.queries(3, 512), rows[5:6]])
    s1, r1 = ix.search(torch.from_numpy(qs).cuda(), 10)
    for g in (2, 4, 8):
        bounds = np.linspace(0, rows.shape[0], g + 1).astype(np.int64)
        parts_s, parts_r = [], []
        for i in range(g):
            shard = mmr.ResidentIndex(ix.rows[bounds[i]:bounds[i + 1]], row_base=int(bounds[i]))
            s, r = shard.search(torch.from_numpy(qs).cuda(), 10)
            parts_s.append(s)
            parts_r.append(r)
        ms, mr = mmr.merge_topk(torch.stack(parts_s), torch.stack(parts_r))
        assert torch.equal(mr, r1), f"G={g}"
        assert torch.equal(ms, s1), f"G={g}"
    ix.close()


def test_zero_query_and_error_paths(mmr):
    rows = util.unit_rows(1000, 512, seed=71)
    ix = mmr.ResidentIndex.from_f32(rows, dtype="bf16")
    s, r = ix.search_host(np.zeros((1, 512), np.float32), 4)
    assert r[0].tolist() == [0, 1, 2, 3] and (s[0] == 0).all()   # cos := 0 for a zero query; ties -> row asc
    with pytest.raises(mmr.NativeError):
        ix.search_host(util.queries(1, 512), 65)                 # k > MMR_MAX_K
    with pytest.raises(mmr.NativeError):
        ix.search_host(util.queries(1, 512), 4, [3])             # segment out of range
    with pytest.raises(ValueError):
        ix.search_host(util.queries(1, 384), 4)
    ix.close()


def test_loader_normalises_like_the_reference(mmr):
    raw = np.random.default_rng(81).standard_normal((5000, 384)).astype(np.float32) * 3
    raw[17] = 0
    want = np.stack([np.asarray(ofs.normalize(v), dtype=np.float32) for v in raw])
    ix = mmr.ResidentIndex.from_f32(raw, dtype="f32", normalize=True)
    got = ix.rows.cpu().numpy()
    assert np.abs(got - want).max() < 2e-7 and (got[17] == 0).all()
    ix2 = mmr.ResidentIndex.from_f32(torch.from_numpy(raw).cuda(), dtype="bf16", normalize=True)
    assert np.abs(ix2.rows.float().cpu().numpy() - ofs.bf16_round(want)).max() < 2 ** -8
    ix.close(); ix2.close()


def test_pipelined_launches_give_identical_results(mmr, monkeypatch):
    """MMR_PDL=1: back-to-back searches overlap (programmatic dependent launch); every result must equal the
    serialized launch's."""
    rows = util.unit_rows(400_000, 512, seed=101)
    ix = mmr.ResidentIndex.from_f32(rows, dtype="bf16")
    qs = torch.from_numpy(util.queries(64, 512)).cuda()
    torch.cuda.synchronize()
    results = {}
    for mode in ("0", "1"):
        mmr._native.set_option("MMR_PDL", mode)
        outs = [(torch.empty((1, 10), dtype=torch.float32, device="cuda"), torch.empty((1, 10), dtype=torch.int64, device="cuda"))
                for _ in range(64)]
        for rep in range(3):
            for i in range(64):
                ix.search(qs[i:i + 1], 10, out=outs[i])
        torch.cuda.synchronize()
        results[mode] = (torch.cat([o[0] for o in outs]).cpu(), torch.cat([o[1] for o in outs]).cpu())
    mmr._native.set_option("MMR_PDL", None)
    assert torch.equal(results["0"][1], results["1"][1])
    assert torch.equal(results["0"][0], results["1"][0])
    for i in (0, 17, 63):
        util.check_topk(results["1"][0][i].numpy(), results["1"][1][i].numpy(), util.oracle_scores(rows, qs[i].cpu().numpy()),
                        10, util.TOL_BF16, what=f"pdl q{i}")
    ix.close()
