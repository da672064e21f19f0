"""Batched metadata join: the chunk lookups behind retrieve_text / retrieve_images.

The reference resolves every hit with its own `SELECT * FROM chunks WHERE id = ?` (`MetadataStore.get_chunk`,
app/storage/schema.py:203-214, called in the loops at app/ml/retrieve.py:55-67 and :86-98): up to 62 statements per
request, which dominates the request once the scan takes about a millisecond.  Two drop-ins for
`app.ml.retrieve._METADATA_STORE`, both offering the reference's `get_chunk(id)` plus `get_chunks(ids) -> {id: chunk}`
(which `retrieve._join` uses when present):

  * BatchedMetadataStore   the reference's own SQLite file, one `WHERE id IN (...)` statement per result list
  * ColumnarChunkTable     a host-resident columnar copy (Arrow columns + the sorted 64-bit hash index of hosttable.py),
                           no SQL on the request path at all; `from_sqlite` loads the reference's `chunks` table

Chunks come back as light objects with the attributes `retrieve_*` reads (`id, document_id, modality, text, page_no,
start_ts, end_ts, file_path, meta`), the same names as the reference's pydantic `Chunk` (schema.py:32-45).
"""
from __future__ import annotations

import json
import sqlite3
from types import SimpleNamespace
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np

CHUNK_FIELDS = ("id", "document_id", "modality", "text", "page_no", "start_ts", "end_ts", "file_path", "meta")


def _chunk(row: Dict[str, Any]) -> SimpleNamespace:
    meta = row.get("meta")
    if isinstance(meta, str) or meta is None:
        meta = json.loads(meta or "{}")
    return SimpleNamespace(id=row["id"], document_id=row.get("document_id"), modality=row.get("modality"),
                           text=row.get("text"), page_no=row.get("page_no"), start_ts=row.get("start_ts"),
                           end_ts=row.get("end_ts"), file_path=row.get("file_path"), meta=meta)


class BatchedMetadataStore:
    """The reference's SQLite metadata file, read with one statement per result list."""

    MAX_VARS = 900  # SQLite's default limit on bound parameters is 999

    def __init__(self, db_path: str) -> None:
        self._conn = sqlite3.connect(db_path, check_same_thread=False)
        self._conn.row_factory = sqlite3.Row
        self.statements = 0

    def get_chunk(self, chunk_id: str):
        return self.get_chunks([chunk_id]).get(chunk_id)

    def get_chunks(self, chunk_ids: Sequence[str]) -> Dict[str, SimpleNamespace]:
        ids = list(dict.fromkeys(chunk_ids))
        out: Dict[str, SimpleNamespace] = {}
        for i in range(0, len(ids), self.MAX_VARS):
            part = ids[i:i + self.MAX_VARS]
            cur = self._conn.execute(f"SELECT * FROM chunks WHERE id IN ({','.join('?' * len(part))})", part)
            self.statements += 1
            for row in cur.fetchall():
                out[row["id"]] = _chunk(dict(row))
        return out

    def close(self) -> None:
        self._conn.close()


class ColumnarChunkTable:
    """Host-resident columnar chunk table: id -> row through a sorted 64-bit hash index, values out of Arrow columns."""

    def __init__(self, table=None) -> None:
        import pyarrow as pa

        self._pa = pa
        self._cols: Dict[str, Any] = {}
        self._hash = np.zeros(0, dtype=np.uint64)
        self._row = np.zeros(0, dtype=np.int64)
        self._n = 0
        self._extra: Dict[str, SimpleNamespace] = {}      # rows upserted after the bulk load (newest wins)
        if table is not None:
            self._load(table)

    @classmethod
    def from_sqlite(cls, db_path: str) -> "ColumnarChunkTable":
        """Load the reference's `chunks` table (schema.py:100-118) once."""
        import pyarrow as pa

        conn = sqlite3.connect(db_path)
        try:
            cur = conn.execute(f"SELECT {', '.join(CHUNK_FIELDS)} FROM chunks")
            rows = cur.fetchall()
        finally:
            conn.close()
        cols = list(zip(*rows)) if rows else [[] for _ in CHUNK_FIELDS]
        types = {"page_no": pa.int64(), "start_ts": pa.float64(), "end_ts": pa.float64()}
        arrays = [pa.array(list(c), type=types.get(f, pa.string())) for f, c in zip(CHUNK_FIELDS, cols)]
        return cls(pa.Table.from_arrays(arrays, names=list(CHUNK_FIELDS)))

    def _load(self, table) -> None:
        from .hosttable import hash_strings

        self._cols = {f: table.column(f).combine_chunks() for f in CHUNK_FIELDS if f in table.column_names}
        self._n = table.num_rows
        h = hash_strings(self._cols["id"])
        order = np.argsort(h, kind="stable")
        self._hash, self._row = h[order], order.astype(np.int64)

    def __len__(self) -> int:
        return self._n + len(self._extra)

    def upsert(self, chunks: Iterable[Any]) -> None:
        """Rows written after the bulk load (the reference's upsert_chunks, schema.py:170-201); newest wins."""
        for c in chunks:
            row = {f: getattr(c, f, None) if not isinstance(c, dict) else c.get(f) for f in CHUNK_FIELDS}
            self._extra[row["id"]] = _chunk(row)

    def _row_of(self, ids: List[str]) -> List[int]:
        from .hosttable import hash_strings

        pa = self._pa
        arr = pa.array(ids, pa.string())
        h = hash_strings(arr)
        lo = np.searchsorted(self._hash, h, side="left")
        hi = np.searchsorted(self._hash, h, side="right")
        out = []
        id_col = self._cols["id"]
        for i, cid in enumerate(ids):
            found = -1
            for j in range(int(lo[i]), int(hi[i])):
                r = int(self._row[j])
                if id_col[r].as_py() == cid:
                    found = max(found, r)          # the reference's ON CONFLICT keeps one row per id; last wins here
            out.append(found)
        return out

    def get_chunks(self, chunk_ids: Sequence[str]) -> Dict[str, SimpleNamespace]:
        ids = list(dict.fromkeys(chunk_ids))
        out: Dict[str, SimpleNamespace] = {}
        rest = []
        for cid in ids:
            hit = self._extra.get(cid)
            if hit is not None:
                out[cid] = hit
            else:
                rest.append(cid)
        if rest and self._n:
            for cid, r in zip(rest, self._row_of(rest)):
                if r >= 0:
                    out[cid] = _chunk({f: col[r].as_py() for f, col in self._cols.items()})
        return out

    def get_chunk(self, chunk_id: str):
        return self.get_chunks([chunk_id]).get(chunk_id)


__all__ = ["BatchedMetadataStore", "ColumnarChunkTable"]
