"""Host mirror of the reference orchestration (package retrieve.py) -- CPU only, no scan.

Reads like the reference's own tests (tests/test_retrieve.py, tests/test_generate.py, tests/test_cache.py):
same doubles injected through the same module seams, plus bit-for-bit checks of the host fusion against
the golden vectors produced by the reference's own function bodies."""
import copy
import importlib
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest

PKG = "multimodal-rag-for-image-text-search_b200"
retrieve = importlib.import_module(PKG + ".retrieve")
cache = importlib.import_module(PKG + ".cache")
settings_mod = importlib.import_module(PKG + ".settings")
versions = importlib.import_module(PKG + ".versions")

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "fusion_golden.json")) as fh:
    FUSION = json.load(fh)


class DummyStore:
    def __init__(self, text_rows, image_rows):
        self._text, self._image = text_rows, image_rows

    def search_text(self, user_id, vec, top_k):
        return self._text[:top_k]

    def search_image(self, user_id, vec, top_k):
        return self._image[:top_k]


class DummyMetadata:
    def __init__(self, chunks):
        self._chunks = chunks

    def get_chunk(self, chunk_id):
        return self._chunks.get(chunk_id)


class DummyCrossEncoder:
    def predict(self, pairs):
        return np.linspace(0.1, 0.9, len(pairs))


def Chunk(id, document_id, modality, text=None, meta=None, **kw):
    return SimpleNamespace(id=id, document_id=document_id, modality=modality, text=text, meta=meta or {},
                           page_no=kw.get("page_no"), start_ts=None, end_ts=None, file_path=kw.get("file_path"))


@pytest.fixture(autouse=True)
def clean(monkeypatch):
    cache.clear_all_caches()
    monkeypatch.setattr(retrieve.settings, "retrieval", settings_mod.RetrievalSettings())
    yield
    cache.clear_all_caches()


def _wire(monkeypatch, store, metadata, cross_encoder, version=1):
    monkeypatch.setattr(retrieve, "_LANCEDB_STORE", store)
    monkeypatch.setattr(retrieve, "_METADATA_STORE", metadata)
    monkeypatch.setattr(retrieve, "embed_text_batch", lambda texts: np.ones((1, 384), dtype=np.float32))
    monkeypatch.setattr(retrieve, "embed_query_for_images", lambda query: np.ones(512, dtype=np.float32))
    monkeypatch.setattr(retrieve, "_get_cross_encoder", lambda: cross_encoder)
    monkeypatch.setattr(retrieve, "get_index_version", lambda user_id: version)


def test_retrieve_fusion_reference_fixture(monkeypatch):
    """reference tests/test_retrieve.py:45-72 -- but asserting what the reference CODE returns (golden),
    not its knife-edge `fused[0] == 't1'`."""
    text_rows = [{"chunk_id": "t1", "score": 0.8, "meta": {}}, {"chunk_id": "t2", "score": 0.6, "meta": {}}]
    image_rows = [{"chunk_id": "i1", "score": 0.7, "meta": {}}]
    chunks = {
        "t1": Chunk("t1", "doc1", "text", text="alpha"),
        "t2": Chunk("t2", "doc2", "text", text="beta"),
        "i1": Chunk("i1", "doc3", "image", meta={"file_path": "/tmp/img.jpg"}),
    }
    _wire(monkeypatch, DummyStore(text_rows, image_rows), DummyMetadata(chunks), DummyCrossEncoder())
    fused = retrieve.retrieve("user", "example query")
    want = FUSION["rerank_fuse"][0]["expect"]
    assert [f["chunk_id"] for f in fused] == [w["chunk_id"] for w in want]
    assert [f["combined_score"] for f in fused] == [w["combined_score"] for w in want]
    assert fused[1]["modality"] == "text" and fused[1]["metadata"]["doc_id"] == "doc1"
    assert fused[0]["metadata"]["file_path"] == "/tmp/img.jpg" and fused[0]["text"] is None


def test_retrieval_cache_invalidation(monkeypatch):
    """reference tests/test_retrieve.py:75-89."""
    _wire(monkeypatch, DummyStore([], []), DummyMetadata({}), False, version=1)
    retrieve.retrieve("user", "question")
    monkeypatch.setattr(retrieve, "get_index_version", lambda user_id: 2)
    assert retrieve.retrieve("user", "question") == []


def test_drop_rules_and_defaults(monkeypatch):
    text_rows = [{"chunk_id": c, "score": s, "meta": {}} for c, s in (("a", .9), ("gone", .8), ("empty", .7), ("b", .6))]
    image_rows = [{"chunk_id": c, "score": s, "meta": {}} for c, s in (("i", .5), ("gone", .4))]
    chunks = {"a": Chunk("a", "d", "text", text="x"), "empty": Chunk("empty", "d", "text", text=""),
              "b": Chunk("b", "d", "text", text="y", page_no=3), "i": Chunk("i", "d", "image")}
    seen = {}

    class Spy(DummyStore):
        def search_text(self, user_id, vec, top_k):
            seen["text"] = (user_id, len(vec), top_k)
            return super().search_text(user_id, vec, top_k)

        def search_image(self, user_id, vec, top_k):
            seen["image"] = (user_id, len(vec), top_k)
            return super().search_image(user_id, vec, top_k)

    _wire(monkeypatch, Spy(text_rows, image_rows), DummyMetadata(chunks), False)
    t = retrieve.retrieve_text("u", "q")
    i = retrieve.retrieve_images("u", "q")
    assert [x["chunk_id"] for x in t] == ["a", "b"] and t[1]["metadata"]["page_no"] == 3
    assert [x["chunk_id"] for x in i] == ["i"] and i[0]["text"] is None
    assert seen["text"] == ("u", 384, 50) and seen["image"] == ("u", 512, 12)  # INDEX_TOPK_TEXT / _IMG defaults


@pytest.mark.parametrize("case", FUSION["z_scores"])
def test_z_scores_golden(case):
    assert retrieve._z_scores(case["values"]) == case["expect"]


@pytest.mark.parametrize("case", FUSION["fuse"])
def test_fuse_golden(case, monkeypatch):
    monkeypatch.setattr(retrieve.settings, "retrieval", settings_mod.RetrievalSettings(final_n=case["final_n"]))
    got = retrieve._fuse_results(copy.deepcopy(case["text"]), copy.deepcopy(case["image"]))
    assert got == case["expect"]


@pytest.mark.parametrize("case", FUSION["rerank_fuse"])
def test_rerank_fuse_golden(case, monkeypatch):
    monkeypatch.setattr(retrieve.settings, "retrieval",
                        settings_mod.RetrievalSettings(final_n=case["final_n"], rerank_topk=case["rerank_topk"]))
    replay = list(case["predict"])
    model = SimpleNamespace(predict=lambda pairs: replay[: len(pairs)]) if replay else False
    monkeypatch.setattr(retrieve, "_get_cross_encoder", lambda: model)
    reranked = retrieve._rerank_text("example query", copy.deepcopy(case["text"]))
    assert reranked == case["reranked"]
    assert retrieve._fuse_results(reranked, copy.deepcopy(case["image"])) == case["expect"]


@pytest.mark.parametrize("case", FUSION["confidence"])
def test_confidence_golden(case, monkeypatch):
    monkeypatch.setattr(retrieve.settings, "retrieval", settings_mod.RetrievalSettings(confidence_tau=case["tau"]))
    assert retrieve._confidence_low(case["items"]) is case["expect"]


def test_generate_low_confidence_gate():
    """reference tests/test_generate.py:18-22: score 0.1 < tau 0.25 -> low confidence."""
    items = [{"modality": "text", "score": 0.1, "metadata": {"doc_id": "doc1"}, "text": "sample"}]
    assert retrieve._confidence_low(items) is True
    assert retrieve._confidence_low([{"score": 1.0, "combined_score": 1.0}]) is False


def test_cache_keys():
    """reference tests/test_cache.py."""
    cache.set_query_embeddings(" test Query ", np.ones(384, np.float32), np.ones(512, np.float32), ttl=1)
    assert cache.get_query_embeddings("test query") is not None
    cache.set_retrieval_results("user", "q", 1, [1])
    assert cache.get_retrieval_results("user", "Q", 1) == [1]
    assert cache.get_retrieval_results("user", "Q", 2) is None


def test_settings_env_names():
    env = {"RERANK_ENABLED": "false", "INDEX_TOPK_TEXT": "10", "INDEX_TOPK_IMG": "7", "RERANK_TOPK": "3",
           "FINAL_N": "2", "CONFIDENCE_TAU": "0.5"}
    s = settings_mod.load_retrieval_settings(env)
    assert s == settings_mod.RetrievalSettings(False, 10, 7, 3, 2, 0.5)
    d = settings_mod.load_retrieval_settings({"INDEX_TOPK_TEXT": "oops"})
    assert d == settings_mod.RetrievalSettings()  # reference defaults: True, 50, 12, 8, 4, 0.25


def test_version_file(tmp_path):
    vf = versions.VersionFile(str(tmp_path / "lance" / "index_versions.json"))
    assert vf.get("u") == 0
    assert vf.bump("u") == 1 and vf.bump("u") == 2 and vf.bump("v") == 1
    other = versions.VersionFile(str(tmp_path / "lance" / "index_versions.json"))  # the reader process
    assert other.get("u") == 2 and other.get("v") == 1 and other.get("w") == 0
    assert json.load(open(tmp_path / "lance" / "index_versions.json")) == {"u": 2, "v": 1}
    mem = versions.VersionFile(None)
    assert mem.bump("x") == 1 and mem.get("x") == 1


def test_batched_metadata_join_is_used_when_offered(monkeypatch):
    text_rows = [{"chunk_id": c, "score": s, "meta": {}} for c, s in (("a", .9), ("gone", .8), ("b", .6))]
    chunks = {"a": Chunk("a", "d", "text", text="x"), "b": Chunk("b", "d", "text", text="y")}
    calls = {"bulk": 0, "single": 0}

    class Meta:
        def get_chunks(self, ids):
            calls["bulk"] += 1
            return {i: chunks[i] for i in ids if i in chunks}

        def get_chunk(self, cid):
            calls["single"] += 1
            return chunks.get(cid)

    _wire(monkeypatch, DummyStore(text_rows, []), Meta(), False)
    got = retrieve.retrieve_text("u", "q")
    assert [g["chunk_id"] for g in got] == ["a", "b"] and calls == {"bulk": 1, "single": 0}


def test_host_fusion_matches_oracle_on_random_requests():
    """Property check: the product's host fusion (_rerank_text + _fuse_results + _confidence_low) against the
    golden-pinned oracle on seeded random requests (ties, empty lists, partial rerank, every FINAL_N)."""
    from oracle import fusion as ofu

    rng = np.random.default_rng(99)
    for trial in range(300):
        nt, ni = int(rng.integers(0, 51)), int(rng.integers(0, 13))
        quant = rng.choice([0, 2, 4])                       # coarse scores -> ties
        def scores(n, lo, hi):
            s = rng.uniform(lo, hi, size=n)
            s = np.round(s, quant) if quant else s
            return sorted((float(np.float32(x)) if trial % 2 else float(x) for x in s), reverse=True)
        text = [{"chunk_id": f"t{j}", "modality": "text", "score": sc, "metadata": {}, "text": "" if rng.random() < 0.1 else f"x{j}"}
                for j, sc in enumerate(scores(nt, -0.2, 0.95))]
        image = [{"chunk_id": f"i{j}", "modality": "image", "score": sc, "metadata": {}, "text": None}
                 for j, sc in enumerate(scores(ni, 0.0, 0.5))]
        final_n, rerank_topk = int(rng.choice([1, 4, 10, 70])), int(rng.choice([0, 3, 8, 60]))
        tau = float(rng.choice([0.0, 0.25, 0.9]))
        logits = [float(x) for x in rng.normal(0, 3, size=64).astype(np.float32)]
        use_rerank = bool(trial % 3)
        cfg = settings_mod.RetrievalSettings(use_rerank=use_rerank, rerank_topk=rerank_topk, final_n=final_n, confidence_tau=tau)
        retrieve.settings.retrieval = cfg
        model = SimpleNamespace(predict=lambda pairs: logits[: len(pairs)])
        old = retrieve._get_cross_encoder
        retrieve._get_cross_encoder = lambda: model
        try:
            mine = retrieve._fuse_results(retrieve._rerank_text("q", copy.deepcopy(text)), copy.deepcopy(image))
            low = retrieve._confidence_low(mine)
        finally:
            retrieve._get_cross_encoder = old
        ref_r = ofu.rerank_text("q", copy.deepcopy(text), lambda pairs: logits[: len(pairs)], use_rerank, rerank_topk)
        ref = ofu.fuse_results(ref_r, copy.deepcopy(image), final_n)
        assert mine == ref, trial
        assert low is ofu.confidence_low(ref, tau)
