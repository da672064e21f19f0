"""K5 fusion/gate bit-exactness, and the B200Store / retrieve() drop-in against the oracle store."""
import copy
import importlib
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import flat_search as ofs
from oracle import fusion as ofu
from tests import util

pytestmark = pytest.mark.gpu
PKG = "multimodal-rag-for-image-text-search_b200"
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mmr():
    pkg = importlib.import_module(PKG)
    pkg._native.lib()
    return pkg


def _oracle_fuse(cos_t, cos_i, final_n, tau):
    """What search_* -> _fuse_results -> _confidence_low give for f32 cosine scores (rerank off)."""
    one = np.float32(1.0)
    text = [{"id": ("t", j), "score": 1.0 - float(one - c)} for j, c in enumerate(cos_t)]
    img = [{"id": ("i", j), "score": 1.0 - float(one - c)} for j, c in enumerate(cos_i)]
    fused = ofu.fuse_results(text, img, final_n)
    return fused, ofu.confidence_low(fused, tau)


@pytest.mark.parametrize("kt,ki", [(50, 12), (10, 10), (1, 1), (7, 0), (0, 12), (64, 64), (8, 3)])
def test_fuse_kernel_bit_exact(mmr, kt, ki):
    rng = np.random.default_rng(kt * 100 + ki)
    b = 33
    ts = np.sort(rng.uniform(-0.1, 0.9, size=(b, max(kt, 1))).astype(np.float32), axis=1)[:, ::-1].copy()
    is_ = np.sort(rng.uniform(0.1, 0.4, size=(b, max(ki, 1))).astype(np.float32), axis=1)[:, ::-1].copy()
    ts[1] = 0.5                                   # std == 0 -> all z == 0 -> stable order decides
    tr = np.tile(np.arange(max(kt, 1), dtype=np.int64), (b, 1)) + 1000
    ir = np.tile(np.arange(max(ki, 1), dtype=np.int64), (b, 1)) + 5000
    nt = rng.integers(0, kt + 1, size=b) if kt else np.zeros(b, int)
    ni = rng.integers(0, ki + 1, size=b) if ki else np.zeros(b, int)
    nt[0], ni[0] = 0, 0                           # no hits at all -> low confidence
    for j in range(b):
        tr[j, nt[j]:] = -1
        ir[j, ni[j]:] = -1
    dev = "cuda"
    text = (torch.from_numpy(ts[:, :kt].copy()).to(dev), torch.from_numpy(tr[:, :kt].copy()).to(dev)) if kt else None
    img = (torch.from_numpy(is_[:, :ki].copy()).to(dev), torch.from_numpy(ir[:, :ki].copy()).to(dev)) if ki else None
    for final_n, tau in ((4, 0.25), (1, 0.0), (10, 0.9)):
        out = {k: v.cpu().numpy() for k, v in mmr.fuse(text, img, final_n, tau).items()}
        for j in range(b):
            fused, low = _oracle_fuse(ts[j, :nt[j]], is_[j, :ni[j]], final_n, tau)
            assert bool(out["low_conf"][j]) is low, (j, final_n, tau)
            for o in range(final_n):
                if o < len(fused):
                    kind, pos = fused[o]["id"]
                    assert out["modality"][j, o] == (0 if kind == "t" else 1)
                    assert out["rows"][j, o] == (tr[j, pos] if kind == "t" else ir[j, pos])
                    assert out["combined"][j, o] == fused[o]["combined_score"], (j, o)   # bit-exact float64
                    assert out["score"][j, o] == fused[o]["score"]
                else:
                    assert out["rows"][j, o] == -1 and out["modality"][j, o] == -1


def _rows(n, dim, seed, users, prefix):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, dim)).astype(np.float32) * 2
    return [SimpleNamespace(chunk_id=f"{prefix}{i}", user_id=users[i % len(users)], document_id=f"d{i % 5}",
                            modality="text" if prefix == "t" else "image", embedding=emb[i].tolist(),
                            meta={"i": i}) for i in range(n)]


def test_store_matches_oracle_store(mmr):
    """search_text / search_image through B200Store vs the oracle's LanceDBStore restatement: same chunk ids
    (tolerance-aware), scores within 1e-3, same dict shape; upsert = delete + add; unknown tenant -> []."""
    users = ["alice", "bob", "o'brien"]
    trows, irows = _rows(3000, 384, 1, users, "t"), _rows(800, 512, 2, users, "i")
    gpu, cpu = mmr.B200Store(), ofs.OracleStore()
    for st in (gpu, cpu):
        st.upsert_text_vectors([mmr.VectorRow(**r.__dict__) for r in trows] if st is gpu else trows)
        st.upsert_image_vectors([mmr.VectorRow(**r.__dict__) for r in irows] if st is gpu else irows)
    moved = copy.deepcopy(trows[10])
    moved.embedding = trows[11].embedding          # re-upsert an existing chunk with a new vector
    gpu.upsert_text_vectors([mmr.VectorRow(**moved.__dict__)])
    cpu.upsert_text_vectors([moved])
    assert gpu.get_index_version("alice") >= 2 and gpu.get_index_version("nobody") == 0
    rng = np.random.default_rng(3)
    for user in users + ["nobody"]:
        for fn, dim, k in (("search_text", 384, 50), ("search_image", 512, 12), ("search_text", 384, 0)):
            q = rng.standard_normal(dim).astype(np.float32)
            got = getattr(gpu, fn)(user, q.tolist(), k)
            want = getattr(cpu, fn)(user, q.tolist(), k)
            assert len(got) == len(want)
            if not want:
                assert got == []
                continue
            assert set(got[0]) == {"chunk_id", "score", "meta"} and isinstance(got[0]["score"], float)
            wscore = {w["chunk_id"]: w["score"] for w in want}
            kth = want[-1]["score"]
            for g in got:
                if g["chunk_id"] in wscore:
                    assert abs(g["score"] - wscore[g["chunk_id"]]) <= util.TOL_BF16
                else:
                    assert g["score"] >= kth - 2 * util.TOL_BF16     # only boundary swaps inside the tolerance
                assert g["meta"] == {"i": int(g["chunk_id"][1:])}
            clear = [w["chunk_id"] for w in want if w["score"] > kth + 2 * util.TOL_BF16]
            assert set(clear) <= {g["chunk_id"] for g in got}
            assert [g["score"] for g in got] == sorted((g["score"] for g in got), reverse=True)
    # the re-upserted chunk now sits next to its twin
    hits = gpu.search_text(moved.user_id, moved.embedding, 5)
    assert {h["chunk_id"] for h in hits[:2]} == {"t10", "t11"} or hits[0]["chunk_id"] == "t10"
    # batched entry point == single calls
    qs = rng.standard_normal((5, 384)).astype(np.float32)
    us = ["alice", "bob", "nobody", "alice", "o'brien"]
    batch = gpu.search_text_batch(us, qs, 10)
    for u, q, b in zip(us, qs, batch):
        assert b == gpu.search_text(u, q.tolist(), 10)


def test_arrow_load_and_retrieve_end_to_end(mmr):
    """Reference schema in (Arrow), retrieve()/retrieve_text()/retrieve_images() out, fused like the oracle."""
    retrieve = importlib.import_module(PKG + ".retrieve")
    cache = importlib.import_module(PKG + ".cache")
    settings_mod = importlib.import_module(PKG + ".settings")
    cache.clear_all_caches()
    n_t, n_i = 4000, 1500
    temb, iemb = util.unit_rows(n_t, 384, 5), util.unit_rows(n_i, 512, 6)
    users_t = [f"u{i % 3}" for i in range(n_t)]
    users_i = [f"u{i % 3}" for i in range(n_i)]
    store = mmr.B200Store()
    store.load_arrow("text_collection", mmr.make_arrow_table([f"t{i}" for i in range(n_t)], users_t, ["d"] * n_t,
                                                             ["text"] * n_t, temb, [json.dumps({"i": i}) for i in range(n_t)]))
    store.load_arrow("image_collection", mmr.make_arrow_table([f"i{i}" for i in range(n_i)], users_i, ["d"] * n_i,
                                                              ["image"] * n_i, iemb, [None] * n_i))
    chunks = {f"t{i}": SimpleNamespace(id=f"t{i}", document_id="d", modality="text", text=f"text {i}", meta={},
                                       page_no=i, start_ts=None, end_ts=None, file_path=None) for i in range(n_t)}
    chunks.update({f"i{i}": SimpleNamespace(id=f"i{i}", document_id="d", modality="image", text=None, meta={},
                                            page_no=None, start_ts=None, end_ts=None, file_path=f"/f/{i}.jpg") for i in range(n_i)})
    qt, qi = util.queries(1, 384, seed=9)[0], util.queries(1, 512, seed=10)[0]
    retrieve.configure(store=store, metadata=SimpleNamespace(get_chunk=chunks.get),
                       text_encoder=lambda texts: qt[None, :], image_query_encoder=lambda q: qi,
                       retrieval_settings=settings_mod.RetrievalSettings(use_rerank=False))
    fused = retrieve.retrieve("u1", "what is in the picture?")
    # oracle: same tenant rows, fp32
    sel_t = np.array([i for i in range(n_t) if i % 3 == 1]); sel_i = np.array([i for i in range(n_i) if i % 3 == 1])
    dt, it = ofs.flat_search(temb[sel_t], qt, 50); di, ii = ofs.flat_search(iemb[sel_i], qi, 12)
    text = [{"chunk_id": f"t{sel_t[j]}", "score": 1.0 - float(d)} for d, j in zip(dt, it)]
    img = [{"chunk_id": f"i{sel_i[j]}", "score": 1.0 - float(d)} for d, j in zip(di, ii)]
    want = ofu.fuse_results(text, img, 4)
    assert len(fused) == 4 and all("combined_score" in f and "metadata" in f for f in fused)
    assert [f["chunk_id"] for f in fused[:2]] == [w["chunk_id"] for w in want[:2]]
    for f, w in zip(fused, want):
        assert abs(f["combined_score"] - w["combined_score"]) < 0.05   # z-scores amplify the 1e-3 score tolerance
    assert retrieve._confidence_low(fused) is ofu.confidence_low(want, 0.25)
    assert retrieve.retrieve("u1", "what is in the picture?") is fused              # cache hit, same version
    assert retrieve.retrieve("nobody", "x") == []
    # device fusion of the same request agrees with the host fusion bit for bit
    ts, tr = store._text_table.resident().search_ranges(torch.from_numpy(qt[None]).cuda(), 50, [store._text_table._ranges["u1"]])
    is_, ir = store._image_table.resident().search_ranges(torch.from_numpy(qi[None]).cuda(), 12, [store._image_table._ranges["u1"]])
    out = mmr.fuse((ts, tr), (is_, ir), 4, 0.25)
    assert out["combined"][0].cpu().tolist() == [f["combined_score"] for f in fused]
    assert bool(out["low_conf"][0]) is retrieve._confidence_low(fused)
    cache.clear_all_caches()


def test_incremental_upserts_track_the_oracle_store(mmr):
    """SURVEY 8f rank 1 -- index refresh without re-uploading the table: upserts append per-tenant delta segments and
    tombstone replaced rows (NaN rows are never returned); results keep matching the oracle store after every step and
    compactions stay rare."""
    rng = np.random.default_rng(5)
    users = ["a", "b", "c", "d"]
    gpu, cpu = mmr.B200Store(), ofs.OracleStore()

    def make(ids, user_of):
        return [SimpleNamespace(chunk_id=f"t{i}", user_id=user_of(i), document_id="d", modality="text",
                                embedding=rng.standard_normal(384).astype(np.float32).tolist(), meta={"i": int(i)}) for i in ids]

    def check(tag):
        for u in users + ["nobody"]:
            q = rng.standard_normal(384).astype(np.float32)
            got, want = gpu.search_text(u, q.tolist(), 10), cpu.search_text(u, q.tolist(), 10)
            assert len(got) == len(want), (tag, u)
            if want:
                kth = want[-1]["score"]
                ws = {w["chunk_id"]: w["score"] for w in want}
                for g in got:
                    assert (g["chunk_id"] in ws and abs(g["score"] - ws[g["chunk_id"]]) <= util.TOL_BF16) or \
                        g["score"] >= kth - 2 * util.TOL_BF16, (tag, u, g)
                assert {w["chunk_id"] for w in want if w["score"] > kth + 2 * util.TOL_BF16} <= {g["chunk_id"] for g in got}
        qs = rng.standard_normal((6, 384)).astype(np.float32)
        us = ["a", "c", "b", "a", "nobody", "d"]
        batch = gpu.search_text_batch(us, qs, 5)             # mixed tenants, several ranges each -> one ranges launch
        for u, q, b in zip(us, qs, batch):
            assert [h["chunk_id"] for h in b] == [h["chunk_id"] for h in gpu.search_text(u, q.tolist(), 5)]

    first = make(range(6000), lambda i: users[i % 4])
    for st in (gpu, cpu):
        st.upsert_text_vectors([mmr.VectorRow(**r.__dict__) for r in first] if st is gpu else first)
    check("initial")
    assert gpu._text_table.rebuilds == 1
    next_id = 6000
    for step in range(6):
        new = make(range(next_id, next_id + 40), lambda i: users[(i * 7) % 4])
        next_id += 40
        over = make(rng.choice(next_id - 40, size=25, replace=False), lambda i: users[i % 4])   # replace existing chunks
        batch = new + over
        old_vec = None
        if step == 2:
            victim = over[0]
            hr = gpu._text_table.host_row_of(victim.chunk_id)
            old_vec = gpu._text_table._host_rows(np.array([hr]))[0].copy()
        for st in (gpu, cpu):
            st.upsert_text_vectors([mmr.VectorRow(**r.__dict__) for r in batch] if st is gpu else batch)
        check(f"step{step}")
        if old_vec is not None:
            hits = gpu.search_text(victim.user_id, old_vec.tolist(), 3)    # the replaced vector is gone (tombstone)
            assert all(not (h["chunk_id"] == victim.chunk_id and h["score"] > 0.99) for h in hits)
    coll = gpu._text_table
    assert coll.rebuilds == 1 and coll.appends == 6, (coll.rebuilds, coll.appends)
    assert coll._tomb == 6 * 25 and max(len(r) for r in coll._ranges.values()) <= 7
    # enough deltas force a compaction; answers stay the same
    for step in range(3):
        batch = make(range(next_id, next_id + 8), lambda i: users[i % 4])
        next_id += 8
        for st in (gpu, cpu):
            st.upsert_text_vectors([mmr.VectorRow(**r.__dict__) for r in batch] if st is gpu else batch)
        check(f"compaction{step}")
    assert coll.rebuilds >= 2 and coll._tomb == 0
    assert gpu.get_index_version("a") >= 10


def test_retrieve_batch_device_equals_host_retrieve(mmr):
    """B concurrent requests: device path (two batched scans + K5) == per-request host path (retrieve + gate),
    bit for bit on combined scores, same winners, same gate; a winner that fails the join falls back to the host path."""
    retrieve = importlib.import_module(PKG + ".retrieve")
    cache = importlib.import_module(PKG + ".cache")
    settings_mod = importlib.import_module(PKG + ".settings")
    cache.clear_all_caches()
    n_t, n_i = 6000, 2500
    temb, iemb = util.unit_rows(n_t, 384, 15), util.unit_rows(n_i, 512, 16)
    users = ["u0", "u1", "u2"]
    store = mmr.B200Store()
    store.load_arrow("text_collection", mmr.make_arrow_table([f"t{i}" for i in range(n_t)], [users[i % 3] for i in range(n_t)],
                                                             ["d"] * n_t, ["text"] * n_t, temb, ["{}"] * n_t))
    store.load_arrow("image_collection", mmr.make_arrow_table([f"i{i}" for i in range(n_i)], [users[i % 3] for i in range(n_i)],
                                                              ["d"] * n_i, ["image"] * n_i, iemb, ["{}"] * n_i))
    chunks = {f"t{i}": SimpleNamespace(id=f"t{i}", document_id="d", modality="text", text=f"text {i}", meta={},
                                       page_no=i, start_ts=None, end_ts=None, file_path=None) for i in range(n_t)}
    chunks.update({f"i{i}": SimpleNamespace(id=f"i{i}", document_id="d", modality="image", text=None, meta={},
                                            page_no=None, start_ts=None, end_ts=None, file_path=f"/f/{i}.jpg") for i in range(n_i)})
    qtext = {f"query {j}": util.queries(1, 384, seed=100 + j)[0] for j in range(9)}
    qimg = {f"query {j}": util.queries(1, 512, seed=200 + j)[0] for j in range(9)}
    retrieve.configure(store=store, metadata=SimpleNamespace(get_chunk=chunks.get),
                       text_encoder=lambda texts: qtext[texts[0]][None, :], image_query_encoder=lambda q: qimg[q],
                       retrieval_settings=settings_mod.RetrievalSettings(use_rerank=False))
    reqs = [(users[j % 3], f"query {j}") for j in range(9)] + [("nobody", "query 0")]
    dev = retrieve.retrieve_batch_device([u for u, _ in reqs], [q for _, q in reqs])
    for (u, q), (items, low) in zip(reqs, dev):
        cache.clear_all_caches()
        host = retrieve.retrieve(u, q)
        assert [it["chunk_id"] for it in items] == [h["chunk_id"] for h in host], (u, q)
        assert [it["combined_score"] for it in items] == [h["combined_score"] for h in host]
        assert [it["score"] for it in items] == [h["score"] for h in host]
        assert [it["metadata"] for it in items] == [h["metadata"] for h in host]
        assert low is retrieve._confidence_low(host)
    # knock out one winner's chunk: the device result cannot be trusted for that request -> host path result
    victim = dev[0][0][0]["chunk_id"]
    gone = dict(chunks)
    del gone[victim]
    retrieve.configure(metadata=SimpleNamespace(get_chunk=gone.get))
    cache.clear_all_caches()
    again = retrieve.retrieve_batch_device([reqs[0][0]], [reqs[0][1]])
    cache.clear_all_caches()
    assert [it["chunk_id"] for it in again[0][0]] == [h["chunk_id"] for h in retrieve.retrieve(*reqs[0])]
    assert victim not in [it["chunk_id"] for it in again[0][0]]
    cache.clear_all_caches()


def _replay_on_device(mmr, text, image, final_n, tau, logits=None):
    """One golden request through mmr_fuse_f64; returns ([(kind, position, combined)], low_conf)."""
    kt, ki = len(text), len(image)
    ts = torch.tensor([[it["score"] for it in text]], dtype=torch.float64, device="cuda") if kt else None
    tc = torch.tensor([kt], dtype=torch.int32, device="cuda") if kt else None
    is_ = torch.tensor([[it["score"] for it in image]], dtype=torch.float64, device="cuda") if ki else None
    ic = torch.tensor([ki], dtype=torch.int32, device="cuda") if ki else None
    rr = rc = None
    if logits and kt:
        padded = list(logits) + [0.0] * (kt - len(logits))
        rr = torch.tensor([padded], dtype=torch.float64, device="cuda")
        rc = torch.tensor([len(logits)], dtype=torch.int32, device="cuda")
    if not kt and not ki:
        return [], True
    out = mmr.fuse_f64(ts, tc, is_, ic, final_n, tau, rerank=rr, rerank_count=rc)
    comb, idx = out["combined"][0].cpu().tolist(), out["index"][0].cpu().tolist()
    items = [(("t", i) if i < kt else ("i", i - kt)) + (c,) for c, i in zip(comb, idx) if i >= 0]
    return items, bool(out["low_conf"][0].item())


def test_fuse_kernel_reproduces_reference_outputs_on_golden_vectors(mmr):
    """mmr_fuse_f64 against what the REFERENCE'S OWN _rerank_text / _fuse_results / _confidence_low returned when
    oracle/make_golden.py executed them: same winners in the same order, bit-identical float64 combined scores."""
    with open(os.path.join(HERE, "golden", "fusion_golden.json")) as fh:
        golden = json.load(fh)
    checked = 0
    for case in golden["fuse"]:
        text, image = case["text"], case["image"]
        got, low = _replay_on_device(mmr, text, image, case["final_n"], 0.25)
        want = case["expect"]
        pos_t = {it["chunk_id"]: j for j, it in enumerate(text)}
        pos_i = {it["chunk_id"]: j for j, it in enumerate(image)}
        assert len(got) == len(want)
        for (kind, pos, comb), w in zip(got, want):
            assert (kind, pos) == (("t", pos_t[w["chunk_id"]]) if w["modality"] == "text" else ("i", pos_i[w["chunk_id"]]))
            assert comb == w["combined_score"]
        assert low is ofu.confidence_low(want, 0.25)
        checked += 1
    for case in golden["rerank_fuse"]:
        text, image = case["text"], case["image"]
        # the logits went to the first len(predict) of the first rerank_topk items (zip in _rerank_text)
        got, low = _replay_on_device(mmr, text, image, case["final_n"], 0.25, logits=case["predict"])
        want = case["expect"]
        pos_t = {it["chunk_id"]: j for j, it in enumerate(text)}
        pos_i = {it["chunk_id"]: j for j, it in enumerate(image)}
        assert len(got) == len(want), case["rerank_topk"]
        for (kind, pos, comb), w in zip(got, want):
            assert (kind, pos) == (("t", pos_t[w["chunk_id"]]) if w["modality"] == "text" else ("i", pos_i[w["chunk_id"]]))
            assert comb == w["combined_score"]
        assert low is ofu.confidence_low(want, 0.25)
        checked += 1
    assert checked == len(golden["fuse"]) + len(golden["rerank_fuse"]) == 58
    # the reference repository's own fixture (tests/test_retrieve.py:46-52 with the linspace cross encoder)
    first = golden["rerank_fuse"][0]
    got, _ = _replay_on_device(mmr, first["text"], first["image"], 4, 0.25, logits=first["predict"])
    assert [(k, p) for k, p, _ in got] == [("i", 0), ("t", 0), ("t", 1)]


def test_store_fp32_storage_meets_the_1e5_tolerance(mmr):
    """dtype="f32" keeps the reference's stored precision: scores within 1e-5 of the oracle store, same ids."""
    users = ["a", "b"]
    rows = _rows(4000, 512, 21, users, "i")
    gpu, cpu = mmr.B200Store(dtype="f32"), ofs.OracleStore()
    gpu.upsert_image_vectors([mmr.VectorRow(**r.__dict__) for r in rows])
    cpu.upsert_image_vectors(rows)
    rng = np.random.default_rng(22)
    for u in users:
        for _ in range(5):
            q = rng.standard_normal(512).astype(np.float32)
            got, want = gpu.search_image(u, q.tolist(), 12), cpu.search_image(u, q.tolist(), 12)
            assert [g["chunk_id"] for g in got] == [w["chunk_id"] for w in want]
            assert max(abs(g["score"] - w["score"]) for g, w in zip(got, want)) <= util.TOL_F32


def test_micro_batched_requests_equal_per_request_retrieve(mmr):
    """48 request threads -> MicroBatcher -> retrieve_batch_device (batched scans + K5): every thread gets exactly what
    the per-request host path returns, and the requests shared launches."""
    import threading

    retrieve = importlib.import_module(PKG + ".retrieve")
    cache = importlib.import_module(PKG + ".cache")
    settings_mod = importlib.import_module(PKG + ".settings")
    cache.clear_all_caches()
    n_t, n_i = 5000, 2000
    temb, iemb = util.unit_rows(n_t, 384, 25), util.unit_rows(n_i, 512, 26)
    users = [f"u{j}" for j in range(6)]
    store = mmr.B200Store()
    store.load_arrow("text_collection", mmr.make_arrow_table([f"t{i}" for i in range(n_t)], [users[i % 6] for i in range(n_t)],
                                                             ["d"] * n_t, ["text"] * n_t, temb, ["{}"] * n_t))
    store.load_arrow("image_collection", mmr.make_arrow_table([f"i{i}" for i in range(n_i)], [users[i % 6] for i in range(n_i)],
                                                              ["d"] * n_i, ["image"] * n_i, iemb, ["{}"] * n_i))
    chunks = {f"t{i}": SimpleNamespace(id=f"t{i}", document_id="d", modality="text", text=f"text {i}", meta={},
                                       page_no=i, start_ts=None, end_ts=None, file_path=None) for i in range(n_t)}
    chunks.update({f"i{i}": SimpleNamespace(id=f"i{i}", document_id="d", modality="image", text=None, meta={},
                                            page_no=None, start_ts=None, end_ts=None, file_path=f"/f/{i}.jpg") for i in range(n_i)})
    nq = 48
    qtext = {f"query {j}": util.queries(1, 384, seed=300 + j)[0] for j in range(nq)}
    qimg = {f"query {j}": util.queries(1, 512, seed=400 + j)[0] for j in range(nq)}
    retrieve.configure(store=store, metadata=SimpleNamespace(get_chunk=chunks.get),
                       text_encoder=lambda texts: qtext[texts[0]][None, :], image_query_encoder=lambda q: qimg[q],
                       retrieval_settings=settings_mod.RetrievalSettings(use_rerank=False))
    store.search_text(users[0], qtext["query 0"].tolist(), 1)          # build the resident copies before the threads start
    store.search_image(users[0], qimg["query 0"].tolist(), 1)
    got = {}
    with mmr.MicroBatcher(retrieve.retrieve_batch_device, max_batch=32, max_wait_ms=50.0) as mb:
        def client(j):
            got[j] = mb.retrieve(users[j % 6], f"query {j}", timeout=60)
        threads = [threading.Thread(target=client, args=(j,)) for j in range(nq)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(60)
        assert sum(mb.batches) == nq and len(mb.batches) < nq
    for j in range(nq):
        cache.clear_all_caches()
        host = retrieve.retrieve(users[j % 6], f"query {j}")
        items, low = got[j]
        assert [it["chunk_id"] for it in items] == [h["chunk_id"] for h in host], j
        assert [it["combined_score"] for it in items] == [h["combined_score"] for h in host]
        assert low is retrieve._confidence_low(host)
    cache.clear_all_caches()


def test_persist_and_restart(mmr, tmp_path):
    """Durable state = Arrow IPC files with the reference schema + index_versions.json; a new process (new store on the
    same path) reloads them into HBM and answers identically."""
    db = str(tmp_path / "lance_db")
    rows_t, rows_i = _rows(1500, 384, 31, ["a", "b"], "t"), _rows(700, 512, 32, ["a", "b"], "i")
    s1 = mmr.B200Store(db)
    s1.upsert_text_vectors([mmr.VectorRow(**r.__dict__) for r in rows_t])
    s1.upsert_image_vectors([mmr.VectorRow(**r.__dict__) for r in rows_i])
    s1.upsert_text_vectors([mmr.VectorRow(**rows_t[3].__dict__)])        # an overwrite: only the live copy is persisted
    s1.persist()
    files = sorted(f for f in os.listdir(db) if not f.endswith(".lock"))
    assert files == ["image_collection.g1.base.arrow", "image_collection.manifest.json", "index_versions.json",
                     "text_collection.g1.base.arrow", "text_collection.manifest.json"], files
    s2 = mmr.B200Store(db)
    assert len(s2._text_table) == 1500 and len(s2._image_table) == 700
    assert s2.get_index_version("a") == s1.get_index_version("a") >= 2
    rng = np.random.default_rng(33)
    for u in ("a", "b"):
        qt, qi = rng.standard_normal(384).astype(np.float32), rng.standard_normal(512).astype(np.float32)
        assert s2.search_text(u, qt.tolist(), 20) == s1.search_text(u, qt.tolist(), 20)
        assert s2.search_image(u, qi.tolist(), 12) == s1.search_image(u, qi.tolist(), 12)
    import pyarrow as pa, pyarrow.ipc as ipc
    with pa.memory_map(os.path.join(db, "text_collection.g1.base.arrow"), "r") as src:
        schema = ipc.open_file(src).schema
    assert [f.name for f in schema] == ["chunk_id", "user_id", "document_id", "modality", "embedding", "meta"]
    assert str(schema.field("embedding").type) == "list<item: float>"


def _gpu_writer(db, seed, lo, hi):
    import importlib
    m = importlib.import_module("multimodal-rag-for-image-text-search_b200")
    st = m.B200Store(db)
    rows = _rows(hi - lo, 512, seed, ["alice", "bob"], f"w{seed}_")
    st.upsert_image_vectors([m.VectorRow(**r.__dict__) for r in rows])


def test_writer_process_upsert_is_visible_to_reader_process(mmr, tmp_path):
    """The reference's deployment: a Celery worker upserts (app/tasks.py:108,165), the API process searches
    (app/ml/retrieve.py:21); they meet on disk.  Here: a second PROCESS with its own B200Store writes, this process's
    store -- created before -- returns the new rows on its next call and uploads only the delta."""
    import multiprocessing as mp
    db = str(tmp_path / "lance_db")
    reader = mmr.B200Store(db)
    base = _rows(3000, 512, 41, ["alice", "bob"], "b")
    reader.upsert_image_vectors([mmr.VectorRow(**r.__dict__) for r in base])
    q = np.random.default_rng(42).standard_normal(512).astype(np.float32)
    before = reader.search_image("alice", q.tolist(), 12)
    assert len(before) == 12 and reader._image_table.rebuilds == 1
    ctx = mp.get_context("spawn")
    p = ctx.Process(target=_gpu_writer, args=(db, 43, 0, 400))
    p.start(); p.join(300)
    assert p.exitcode == 0
    v0 = reader.get_index_version("alice")
    after = reader.search_image("alice", q.tolist(), 12)
    assert reader._image_table.rebuilds == 1 and reader._image_table.appends == 1      # delta only, no re-upload
    assert len(reader._image_table) == 3400 and v0 == 2
    fresh = mmr.B200Store(db)
    assert fresh.search_image("alice", q.tolist(), 12) == after
    # the planted neighbour written by the other process comes back first
    new_rows = _rows(400, 512, 43, ["alice", "bob"], "w43_")
    target = next(r for r in new_rows if r.user_id == "alice")
    top = reader.search_image("alice", list(target.embedding), 1)
    assert top[0]["chunk_id"] == target.chunk_id and top[0]["score"] > 0.99
