"""The flat-scan oracle itself (parity unpinned against LanceDB -- see oracle/__init__.py): internal
consistency, the tie rule, tenant ranges, and the bf16 rounding helper."""
import numpy as np
import pytest
import torch

from oracle import flat_search as ofs
from tests import util


def test_distances_match_float64_bruteforce():
    rows = util.unit_rows(2000, 384, 1)
    q = util.queries(1, 384)[0] * 3.0  # not unit: search re-normalises
    d, ids = ofs.flat_search(rows, q, 10)
    ref = 1.0 - rows.astype(np.float64) @ (q / np.linalg.norm(q)).astype(np.float64)
    order = np.argsort(ref, kind="stable")[:10]
    assert ids.tolist() == order.tolist()
    assert np.abs(d - ref[order]).max() < 1e-6
    d2, ids2 = ofs.flat_search(rows, q, 10, unit_rows=False)
    assert ids2.tolist() == ids.tolist() and np.abs(d2 - d).max() < 1e-6


def test_tie_rule_distance_then_ordinal():
    rows = util.unit_rows(64, 384, 2)
    rows[40] = rows[7]
    rows[3] = rows[7]
    q = rows[7]
    d, ids = ofs.flat_search(rows, q, 4)
    assert ids[:3].tolist() == [3, 7, 40]
    assert d[0] == d[1] == d[2]


def test_k_clamp_ranges_and_empty():
    rows = util.unit_rows(100, 384, 3)
    q = util.queries(1, 384)[0]
    d, ids = ofs.flat_search(rows, q, 0)          # max(top_k, 1)
    assert len(ids) == 1
    d, ids = ofs.flat_search(rows, q, 10, lo=20, hi=25)
    assert len(ids) == 5 and ((ids >= 20) & (ids < 25)).all()
    d, ids = ofs.flat_search(rows, q, 10, lo=30, hi=30)
    assert len(ids) == 0


def test_batch_equals_single():
    rows = util.unit_rows(5000, 512, 4)
    qs = util.queries(5, 512)
    bd, bi = ofs.flat_search_batch(rows, qs, 12, lo=100, hi=4100, block=1000)
    for j in range(5):
        d, ids = ofs.flat_search(rows, qs[j], 12, lo=100, hi=4100)
        assert bi[j].tolist() == ids.tolist()
        assert np.abs(bd[j] - d).max() < 1e-6


def test_bf16_round_matches_torch():
    x = np.random.default_rng(5).standard_normal(4096).astype(np.float32)
    x[:4] = [0.0, -0.0, 1.0, 1.00390625]  # exact tie to even
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert (ofs.bf16_round(x) == want).all()


def test_oracle_store_roundtrip():
    from types import SimpleNamespace as Row
    store = ofs.OracleStore()
    rng = np.random.default_rng(6)
    rows = [Row(chunk_id=f"c{i}", user_id="u1" if i % 2 else "u2", document_id="d", modality="text",
                embedding=rng.standard_normal(384).tolist(), meta={"i": i}) for i in range(40)]
    store.upsert_text_vectors(rows)
    store.upsert_text_vectors([rows[5]])  # delete + add keeps one copy
    hits = store.search_text("u1", rows[5].embedding, 3)
    assert hits[0]["chunk_id"] == "c5" and abs(hits[0]["score"] - 1.0) < 1e-6
    assert all(h["meta"]["i"] % 2 == 1 for h in hits)
    assert store.search_text("nobody", rows[5].embedding, 3) == []
    assert len(store.search_text("u1", rows[5].embedding, 0)) == 1


def test_oracle_agrees_with_an_independent_bruteforce_cosine_knn():
    """Second opinion on the scan oracle (still not LanceDB, which is not installable here): scikit-learn's exact
    brute-force cosine KNN -- an unrelated implementation of "d = 1 - cos, k smallest" -- must return the same ids
    (tie-aware) and distances within fp32 rounding, on unit and non-unit rows, inside a tenant range, for k = 10 and 50."""
    sk = pytest.importorskip("sklearn.neighbors")
    rng = np.random.default_rng(12)
    for dim, n, unit in ((384, 4000, True), (512, 3000, False)):
        rows = util.unit_rows(n, dim, 20 + dim)
        if not unit:
            rows = (rows * rng.uniform(0.5, 2.0, size=(n, 1))).astype(np.float32)
        qs = util.queries(6, dim, seed=30 + dim) * 2.5
        lo, hi = 500, n - 700
        nn = sk.NearestNeighbors(metric="cosine", algorithm="brute").fit(rows[lo:hi].astype(np.float64))
        for k in (10, 50):
            dist, idx = nn.kneighbors(qs.astype(np.float64), n_neighbors=k)
            for j in range(len(qs)):
                d, ids = ofs.flat_search(rows, qs[j], k, lo=lo, hi=hi, unit_rows=unit)
                assert np.abs(d - dist[j]).max() < 2e-6
                want = idx[j] + lo
                for p in range(k):
                    if ids[p] != want[p]:      # only inside a run of (near-)equal distances
                        assert abs(dist[j][p] - d[p]) < 2e-6 and ids[p] in want and want[p] in ids, (dim, k, j, p)
