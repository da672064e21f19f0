"""Batched metadata join: one statement / one columnar lookup per result list instead of the
reference's SELECT per hit (app/ml/retrieve.py:55-67, app/storage/schema.py:203-214)."""
import importlib
import json
import sqlite3
from types import SimpleNamespace

import numpy as np
import pytest

PKG = "multimodal-rag-for-image-text-search_b200"
ct = importlib.import_module(PKG + ".chunktable")
retrieve = importlib.import_module(PKG + ".retrieve")
cache = importlib.import_module(PKG + ".cache")
settings_mod = importlib.import_module(PKG + ".settings")


@pytest.fixture
def sqlite_db(tmp_path):
    """The reference's `chunks` table (app/storage/schema.py:100-118)."""
    path = str(tmp_path / "metadata.sqlite3")
    conn = sqlite3.connect(path)
    conn.execute("""CREATE TABLE chunks (id TEXT PRIMARY KEY, document_id TEXT NOT NULL, modality TEXT NOT NULL, text TEXT,
                    page_no INTEGER, start_ts REAL, end_ts REAL, file_path TEXT, meta TEXT, created_at TEXT, updated_at TEXT)""")
    rows = [(f"c{i}", f"d{i % 3}", "text" if i % 4 else "image", f"passage {i}" if i % 4 else None, i, None, None,
             None if i % 4 else f"/f/{i}.jpg", json.dumps({"i": i}), "t", "t") for i in range(500)]
    rows.append(("o'brien", "d0", "text", "", 1, 0.5, 1.5, None, None, "t", "t"))      # empty text, NULL meta, quote in id
    conn.executemany("INSERT INTO chunks VALUES (?,?,?,?,?,?,?,?,?,?,?)", rows)
    conn.commit()
    conn.close()
    return path


@pytest.mark.parametrize("kind", ["sqlite", "columnar"])
def test_get_chunks_equals_per_id_lookups(sqlite_db, kind):
    store = ct.BatchedMetadataStore(sqlite_db) if kind == "sqlite" else ct.ColumnarChunkTable.from_sqlite(sqlite_db)
    ids = ["c7", "c0", "missing", "c499", "o'brien", "c7"]
    got = store.get_chunks(ids)
    assert set(got) == {"c7", "c0", "c499", "o'brien"}
    c7 = got["c7"]
    assert (c7.id, c7.document_id, c7.modality, c7.text, c7.page_no, c7.meta) == ("c7", "d1", "text", "passage 7", 7, {"i": 7})
    assert got["c0"].modality == "image" and got["c0"].text is None and got["c0"].file_path == "/f/0.jpg"
    assert got["o'brien"].meta == {} and got["o'brien"].text == "" and got["o'brien"].start_ts == 0.5
    assert store.get_chunk("c13").text == "passage 13" and store.get_chunk("nope") is None
    big = store.get_chunks([f"c{i}" for i in range(500)] * 3)                       # more ids than SQLite's variable limit
    assert len(big) == 500


def test_columnar_table_upsert_newest_wins(sqlite_db):
    table = ct.ColumnarChunkTable.from_sqlite(sqlite_db)
    assert len(table) == 501
    table.upsert([SimpleNamespace(id="c7", document_id="d9", modality="text", text="rewritten", page_no=1, start_ts=None,
                                  end_ts=None, file_path=None, meta={"v": 2}),
                  {"id": "new", "document_id": "d9", "modality": "image", "text": None, "file_path": "/n.jpg", "meta": "{}"}])
    got = table.get_chunks(["c7", "new", "c9"])
    assert got["c7"].text == "rewritten" and got["c7"].meta == {"v": 2} and got["new"].file_path == "/n.jpg"
    assert got["c9"].text == "passage 9"


def test_retrieve_text_issues_one_statement_per_result_list(sqlite_db):
    """retrieve_text / retrieve_images through the host mirror: 50 hits -> ONE `WHERE id IN (...)`, same dicts and drop
    rules (missing chunk, empty text) as the per-hit loop of the reference."""
    meta = ct.BatchedMetadataStore(sqlite_db)
    hits = [{"chunk_id": f"c{i}", "score": 1.0 - i * 0.01, "meta": {}} for i in range(1, 40)] + \
           [{"chunk_id": "gone", "score": 0.5, "meta": {}}, {"chunk_id": "o'brien", "score": 0.4, "meta": {}}]
    store = SimpleNamespace(search_text=lambda u, v, k: list(hits), search_image=lambda u, v, k: list(hits[:12]),
                            get_index_version=lambda u: 1)
    cache.clear_all_caches()
    retrieve.configure(store=store, metadata=meta, text_encoder=lambda t: np.ones((1, 384), np.float32),
                       image_query_encoder=lambda q: np.ones(512, np.float32),
                       retrieval_settings=settings_mod.RetrievalSettings(use_rerank=False))
    text = retrieve.retrieve_text("u", "q")
    assert meta.statements == 1
    want = [h for h in hits if h["chunk_id"].startswith("c") and int(h["chunk_id"][1:]) % 4]     # images have no text
    assert [t["chunk_id"] for t in text] == [h["chunk_id"] for h in want]                        # "gone" / empty text dropped
    assert text[0]["metadata"]["doc_id"] == "d1" and text[0]["metadata"]["page_no"] == 1 and text[0]["text"] == "passage 1"
    images = retrieve.retrieve_images("u", "q")
    assert meta.statements == 2 and len(images) == 12 and images[0]["text"] is None
    cache.clear_all_caches()
