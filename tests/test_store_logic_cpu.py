"""Host logic of B200Store (collections, tenant ranges, tombstones, compaction, versions, persistence) on CPU.

The scan itself only exists on the GPU, so ResidentIndex is replaced here by a numpy stand-in with the same interface
(test infrastructure, like the reference's DummyStore); everything above it -- the code that decides WHICH rows a tenant
owns after any sequence of upserts -- is the product code under test."""
import copy
import importlib
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import flat_search as ofs

PKG = "multimodal-rag-for-image-text-search_b200"
store_mod = importlib.import_module(PKG + ".store")


from tests.fake_index import FakeIndex


@pytest.fixture
def cpu_store(monkeypatch, tmp_path):
    monkeypatch.setattr(store_mod, "ResidentIndex", FakeIndex)

    def make(db_path=None):
        st = store_mod.B200Store.__new__(store_mod.B200Store)       # skip the CUDA checks of __init__
        st._init_state(db_path, torch.device("cpu"), "f32")
        return st

    return make


def _mk(rng, ids, user_of, dim=384):
    return [SimpleNamespace(chunk_id=f"t{i}", user_id=user_of(i), document_id="d", modality="text",
                            embedding=rng.standard_normal(dim).astype(np.float32).tolist(), meta={"i": int(i)}) for i in ids]


def _same(got, want, tol=1e-5):
    assert [g["chunk_id"] for g in got] == [w["chunk_id"] for w in want]
    assert all(abs(g["score"] - w["score"]) <= tol and g["meta"] == w["meta"] for g, w in zip(got, want))


def test_upserts_overwrites_and_compaction_track_the_oracle(cpu_store):
    rng = np.random.default_rng(1)
    users = ["a", "b", "c"]
    gpu, cpu = cpu_store(), ofs.OracleStore()
    first = _mk(rng, range(900), lambda i: users[i % 3])
    gpu.upsert_text_vectors([store_mod.VectorRow(**r.__dict__) for r in first])
    cpu.upsert_text_vectors(first)
    next_id = 900
    for step in range(14):
        batch = _mk(rng, range(next_id, next_id + 12), lambda i: users[(i * 5) % 3]) + \
                _mk(rng, rng.choice(next_id, size=9, replace=False), lambda i: users[i % 3])
        if step == 4:
            batch += [copy.deepcopy(batch[0])]                       # the same chunk twice in one batch: last one wins
            batch[-1].embedding = rng.standard_normal(384).astype(np.float32).tolist()
        next_id += 12
        gpu.upsert_text_vectors([store_mod.VectorRow(**r.__dict__) for r in batch])
        if step == 4:
            cpu.upsert_text_vectors(batch[:-1]); cpu.upsert_text_vectors(batch[-1:])
        else:
            cpu.upsert_text_vectors(batch)
        for u in users + ["nobody"]:
            q = rng.standard_normal(384).astype(np.float32)
            _same(gpu.search_text(u, q.tolist(), 10), cpu.search_text(u, q.tolist(), 10))
        qs = rng.standard_normal((4, 384)).astype(np.float32)
        for u, q, b in zip(["a", "c", "nobody", "b"], qs, gpu.search_text_batch(["a", "c", "nobody", "b"], qs, 5)):
            _same(b, cpu.search_text(u, q.tolist(), 5))
    coll = gpu._text_table
    assert coll.appends >= 6 and 2 <= coll.rebuilds <= 6          # deltas most of the time, compaction when ranges pile up
    assert all(len(r) < coll.MAX_RANGES for r in coll._ranges.values())
    assert len(coll) == len({c for c in cpu._text_table.chunk_id})
    assert gpu.get_index_version("a") == 15 and gpu.get_index_version("nobody") == 0
    assert gpu.search_text("a", rng.standard_normal(384).tolist(), 0).__len__() == 1     # max(top_k, 1)
    with pytest.raises(store_mod.N.NativeError):
        gpu.search_text("a", rng.standard_normal(384).tolist(), 65)


def test_empty_store_and_mixed_dims(cpu_store):
    st = cpu_store()
    assert st.search_text("u", [0.0] * 384, 5) == [] and st.search_image_batch(["u"], np.zeros((1, 512)), 5) == [[]]
    st.upsert_text_vectors([])
    st.upsert_image_vectors([store_mod.VectorRow("i1", "u", "d", "image", [1.0] * 512, {})])
    with pytest.raises(ValueError):
        st.upsert_image_vectors([store_mod.VectorRow("i2", "u", "d", "image", [1.0] * 384, {})])
    hit = st.search_image("u", [2.0] * 512, 3)
    assert [h["chunk_id"] for h in hit] == ["i1"] and abs(hit[0]["score"] - 1.0) < 1e-6 and hit[0]["meta"] == {}


def test_persist_and_reload(cpu_store, tmp_path):
    rng = np.random.default_rng(2)
    db = str(tmp_path / "db")
    s1 = cpu_store(db)
    rows = _mk(rng, range(300), lambda i: "a" if i % 2 else "o'brien")
    s1.upsert_text_vectors([store_mod.VectorRow(**r.__dict__) for r in rows])
    s1.upsert_text_vectors([store_mod.VectorRow(**rows[7].__dict__)])
    s1.persist()
    s2 = cpu_store(db)
    assert len(s2._text_table) == 300 and s2.get_index_version("o'brien") == s1.get_index_version("o'brien")
    q = rng.standard_normal(384).astype(np.float32).tolist()
    assert s2.search_text("o'brien", q, 12) == s1.search_text("o'brien", q, 12)
    table = store_mod.make_arrow_table(["x"], ["u"], ["d"], ["text"], np.ones((1, 384), np.float32), [None])
    assert table.schema.names == ["chunk_id", "user_id", "document_id", "modality", "embedding", "meta"]
    s2.load_arrow("text_collection", table)
    assert s2.search_text("u", [1.0] * 384, 1)[0]["chunk_id"] == "x"


def test_query_row_narrows_a_python_list_exactly_like_numpy():
    """_query_row: the reference hands search_* a Python list of floats (retrieve.py:53,84); the C-level narrowing must give
    the float32 values np.asarray(list, float32) gives, for floats, ints, numpy scalars, tuples and arrays, and must fall
    back to numpy for anything array('f') refuses."""
    rng = np.random.default_rng(3)
    vals = (rng.standard_normal(384) * 10.0 ** rng.integers(-6, 6, 384)).tolist() + [0.0, -0.0, 1, -7, 1 / 3, 16777217.0]
    for src in (vals, tuple(vals), [np.float32(v) for v in vals[:50]], np.asarray(vals), np.asarray(vals, dtype=np.float32)):
        got = store_mod._query_row(src)
        want = np.asarray(src, dtype=np.float32)[None, :]
        assert got.dtype == np.float32 and got.shape == want.shape and got.flags.c_contiguous
        assert got.tobytes() == want.tobytes()
    assert store_mod._query_row([]).shape == (1, 0)
    odd = store_mod._query_row([0.1, None, 0.3])          # array('f') refuses None: numpy's answer (nan) through the fallback
    assert odd.shape == (1, 3) and np.isnan(odd[0, 1]) and odd[0, 0] == np.float32(0.1)
