"""HostTable (columnar host master copy) and the durable log: pure host code, no GPU.

Covers what the reference gets from LanceDB on the write side (app/storage/lancedb_store.py:87-101): delete-by-chunk_id
then add, also inside one batch; plus the cross-process contract of the reference's deployment (Celery writer, API
reader: app/tasks.py:108,165, docker-compose.yml:36-45) -- an upsert made durable by one process becomes visible to a
store in another process on its next call."""
import importlib
import multiprocessing as mp
import os
import time
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import flat_search as ofs
from tests.fake_index import make_cpu_store

PKG = "multimodal-rag-for-image-text-search_b200"
ht = importlib.import_module(PKG + ".hosttable")
store_mod = importlib.import_module(PKG + ".store")


def _table(ids, users, dim=8, seed=0, metas=None):
    rng = np.random.default_rng(seed)
    emb = ofs.normalize_rows(rng.standard_normal((len(ids), dim)).astype(np.float32))
    return ht.make_arrow_table(ids, users, ["d"] * len(ids), ["text"] * len(ids), emb, metas or ["{}"] * len(ids)), emb


def test_hash_strings_is_deterministic_and_spreads():
    import pyarrow as pa
    ids = [f"chunk-{i}" for i in range(50_000)] + ["", "a", "ab", "abcdefgh", "abcdefghi", None]
    h = ht.hash_strings(pa.array(ids, pa.string()))
    assert h.dtype == np.uint64 and len(np.unique(h)) == len(ids) - 1        # None hashes like ""
    assert (ht.hash_strings(pa.array(ids, pa.string())) == h).all()
    sl = pa.array(ids, pa.string()).slice(7, 100)                             # sliced arrays carry an offset
    assert (ht.hash_strings(sl) == h[7:107]).all()


def test_append_replace_and_duplicates_inside_a_batch():
    t = ht.HostTable("text_collection")
    tab, emb = _table(["a", "b", "c", "a", "a"], ["u1", "u2", "u1", "u1", "u1"])
    base, n, killed, users = t.append_table(tab)
    assert (base, n, killed.size, users) == (0, 5, 0, ["u1", "u2"])
    assert t.alive[:5].tolist() == [False, True, True, False, True] and len(t) == 3   # the last "a" wins
    assert t.find_alive(["a", "b", "zzz"]).tolist() == [4, 1, -1]
    tab2, _ = _table(["b", "d"], ["u2", "u3"], seed=1)
    base, n, killed, users = t.append_table(tab2)
    assert base == 5 and killed.tolist() == [1] and users == ["u2", "u3"] and len(t) == 4
    assert t.find_alive(["a", "b", "c", "d"]).tolist() == [4, 5, 2, 6]
    assert t.tenants == ["u1", "u2", "u3"] and t.tenant[:7].tolist() == [0, 1, 0, 0, 0, 1, 2]
    assert t.chunk_id_at(6) == "d" and np.allclose(t.gather(np.array([6, 0, 2])), np.stack([_table(["b", "d"], ["u2", "u3"], seed=1)[1][1], emb[0], emb[2]]))
    out = t.to_arrow()
    assert out.column("chunk_id").to_pylist() == ["c", "a", "b", "d"]
    with pytest.raises(ValueError):
        t.append_table(_table(["x"], ["u"], dim=4)[0])


def test_hits_at_reads_ids_and_metas_of_a_result_list_in_one_pass():
    """HostTable.hits_at = what _format_results needs per hit (chunk_id, json.loads(meta) if meta else {}), straight out of
    the Arrow buffers: empty metas recognised on their raw bytes, rows spread over several blocks, sliced tables, and a
    block whose meta column has nulls (slow path through Arrow scalars)."""
    import pyarrow as pa
    t = ht.HostTable("text_collection")
    metas1 = ["{}", "", '{"page": 3, "t": "x y"}', "{}", '{"a": [1, 2]}', "[]"]
    tab1, _ = _table([f"a{i}" for i in range(6)], ["u"] * 6, metas=metas1)
    t.append_table(tab1)
    tab2, _ = _table([f"b{i}" for i in range(40)], ["v"] * 40, seed=1, metas=['{"i": %d}' % i for i in range(40)])
    t.append_table(tab2.slice(5, 20))                                   # a sliced table: the offsets buffer carries an offset
    tab3, _ = _table(["c0", "c1", "c2"], ["w"] * 3, seed=2)
    tab3 = tab3.set_column(tab3.schema.get_field_index("meta"), "meta", pa.array(['{"k": 1}', None, "{}"], pa.string()))
    t.append_table(tab3)
    rows = [2, 0, 1, 5, 4, 3, 6, 25, 10, 26, 27, 28]
    ids, metas = t.hits_at(rows)
    assert ids == ["a2", "a0", "a1", "a5", "a4", "a3", "b5", "b24", "b9", "c0", "c1", "c2"]
    assert metas == [{"page": 3, "t": "x y"}, {}, {}, [], {"a": [1, 2]}, {}, {"i": 5}, {"i": 24}, {"i": 9}, {"k": 1}, {}, {}]
    assert ids == t.values_at("chunk_id", rows)
    assert t.hits_at([]) == ([], [])


def test_index_levels_merge_and_million_row_table_stays_columnar():
    """1M rows in one bulk append, then small upserts: no per-row Python objects, lookups through the sorted hash index."""
    import pyarrow as pa
    n, d = 1_000_000, 4
    emb = np.zeros((n, d), np.float32)
    emb[:, 0] = 1.0
    ids = pa.array(np.arange(n)).cast(pa.string())
    users = pa.array(np.arange(n) % 1000).cast(pa.string())
    const = pa.array(["x"] * 1, pa.string()).take(pa.array(np.zeros(n, np.int32)))
    tab = ht.make_arrow_table(ids, users, const, const, emb, const)
    t = ht.HostTable("image_collection")
    t0 = time.perf_counter()
    t.append_table(tab)
    assert time.perf_counter() - t0 < 20.0
    assert t.blocks[0].emb.base is not None or t.blocks[0].emb.flags["OWNDATA"] is False   # a view, not a copy
    assert len(t) == n and len(t.tenants) == 1000
    t0 = time.perf_counter()
    for step in range(20):                                   # 20 small upserts replacing 5 rows + adding 5 each
        ids2 = [str(step * 7919 + j) for j in range(5)] + [f"new-{step}-{j}" for j in range(5)]
        tab2, _ = _table(ids2, ["7"] * 10, dim=d, seed=step)
        base, m, killed, _ = t.append_table(tab2)
        assert m == 10 and killed.size == 5
    assert time.perf_counter() - t0 < 20.0                   # O(new rows) + one index build, not O(table) per upsert
    assert len(t) == n + 100 and t.find_alive(["0", "new-3-2", "nope"]).tolist()[0] >= n
    assert t.find_alive(["999999"]).tolist() == [999_999]


def _same(got, want, tol=1e-6):
    """The numpy stand-in sums in BLAS order, which depends on where a row sits; ids and metas must match exactly."""
    assert [g["chunk_id"] for g in got] == [w["chunk_id"] for w in want]
    assert all(abs(g["score"] - w["score"]) <= tol and g["meta"] == w["meta"] for g, w in zip(got, want))


def _writer_proc(db, n0, n1):
    st = make_cpu_store(db)
    rng = np.random.default_rng(n0)
    rows = [SimpleNamespace(chunk_id=f"t{i}", user_id="alice", document_id="d", modality="text",
                            embedding=rng.standard_normal(16).astype(np.float32).tolist(), meta={"i": i}) for i in range(n0, n1)]
    st.upsert_text_vectors([store_mod.VectorRow(**r.__dict__) for r in rows])


def test_upsert_in_another_process_becomes_visible(tmp_path):
    db = str(tmp_path / "db")
    reader = make_cpu_store(db)
    q = [1.0] * 16
    assert reader.search_text("alice", q, 5) == [] and reader.get_index_version("alice") == 0
    ctx = mp.get_context("spawn")
    p = ctx.Process(target=_writer_proc, args=(db, 0, 40))
    p.start(); p.join(120)
    assert p.exitcode == 0
    hits = reader.search_text("alice", q, 5)                 # no restart, no explicit reload
    assert len(hits) == 5 and reader.get_index_version("alice") == 1
    assert reader._text_table.rebuilds == 1
    # a second writer process: the reader applies only the new delta file (append, no rebuild) and agrees with a fresh store
    p = ctx.Process(target=_writer_proc, args=(db, 30, 60))  # replaces t30..t39, adds t40..t59
    p.start(); p.join(120)
    assert p.exitcode == 0
    hits2 = reader.search_text("alice", q, 50)
    assert len(hits2) == 50 and reader.get_index_version("alice") == 2
    assert reader._text_table.rebuilds == 1 and reader._text_table.appends == 1 and len(reader._text_table) == 60
    fresh = make_cpu_store(db)
    _same(fresh.search_text("alice", q, 50), hits2)
    # durable compaction by the writer side: the reader notices the new generation and reloads
    fresh.persist()
    files = sorted(os.listdir(db))
    assert not [f for f in files if ".d0" in f] and [f for f in files if f.endswith(".base.arrow")]
    _same(reader.search_text("alice", q, 50), hits2)
    assert reader._text_table._gen == fresh._text_table._gen


def test_concurrent_searches_and_upserts_are_serialised(tmp_path):
    """request threads + a writer thread share one store; the per-collection lock keeps results whole."""
    import threading
    st = make_cpu_store()
    rng = np.random.default_rng(3)
    mk = lambda lo, hi, u: [store_mod.VectorRow(f"{u}{i}", u, "d", "text", rng.standard_normal(16).tolist(), {"u": u}) for i in range(lo, hi)]
    st.upsert_text_vectors(mk(0, 200, "a") + mk(0, 200, "b"))
    errors = []

    def reader(user):
        try:
            for _ in range(60):
                for h in st.search_text(user, rng.standard_normal(16).tolist(), 10):
                    assert h["meta"] == {"u": user} and h["chunk_id"].startswith(user)
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    def writer():
        try:
            for step in range(30):
                st.upsert_text_vectors(mk(step * 3, step * 3 + 6, "a") + mk(200 + step, 201 + step, "b"))
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=reader, args=(u,)) for u in ("a", "b", "a")] + [threading.Thread(target=writer)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    assert len(st._text_table) == 200 + 200 + 30
