"""B200Store -- the duck type of the reference's LanceDBStore (app/storage/lancedb_store.py:24-144) with
the flat cosine scan running on a B200 through libmmr_b200.so.

Same public surface and conventions:
  * VectorRow                                  (lancedb_store.py:12-21)
  * upsert_text_vectors / upsert_image_vectors (:87-101)  delete-by-chunk_id then add, rows L2-normalised
  * search_text / search_image                 (:103-123) -> [{"chunk_id", "score", "meta"}] best first,
    `max(top_k, 1)` hits at most, `[]` for an unknown tenant, errors propagate as exceptions.
  * two shared collections, tenancy = `user_id` (prefilter semantics)

Plus what a resident index needs: `search_*_batch` (micro-batched requests -> one launch), `load_arrow`
(bulk load of a table with the reference's 6-column schema, :33-44), and a per-user version counter
bumped on every upsert (app/ml/index_build.py:33-43) that invalidates the resident copy.

There is no CPU path: constructing a B200Store without the CUDA library / device raises.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .index import ResidentIndex
from .versions import VersionFile


@dataclass
class VectorRow:
    """Payload used when writing vectors (same fields as the reference's VectorRow)."""

    chunk_id: str
    user_id: str
    document_id: str
    modality: str
    embedding: Sequence[float]
    meta: Dict[str, Any]


def _unit_f32(vector: Sequence[float]) -> np.ndarray:
    """LanceDBStore._normalize (:63-69) kept as an f32 array instead of a Python list."""
    arr = np.asarray(vector, dtype=np.float32)
    norm = np.linalg.norm(arr)
    if norm <= 0:
        return arr
    return arr / norm


class _Collection:
    """Host master copy (columns in insertion order) + the tenant-sorted resident copy on the GPU."""

    def __init__(self, name: str, device: torch.device, dtype: str) -> None:
        self.name = name
        self.device = device
        self.dtype = dtype
        self.chunk_id: List[str] = []
        self.user_id: List[str] = []
        self.document_id: List[str] = []
        self.modality: List[str] = []
        self.meta: List[Optional[str]] = []
        self._blocks: List[np.ndarray] = []     # f32 [n_i, D] blocks, concatenation = host row order
        self._alive: List[bool] = []
        self._where: Dict[str, int] = {}        # chunk_id -> host row (alive rows only)
        self._dirty = True
        self._resident: Optional[ResidentIndex] = None
        self._perm = np.empty(0, np.int64)      # resident row -> host row
        self._seg_of: Dict[str, int] = {}
        self.rebuilds = 0

    # -- writes -------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._where)

    def _append(self, chunk_ids, user_ids, doc_ids, modalities, metas, emb: np.ndarray) -> None:
        base = len(self.chunk_id)
        for j, cid in enumerate(chunk_ids):
            old = self._where.get(cid)
            if old is not None:
                self._alive[old] = False          # table.delete("chunk_id == '...'") (:91-92)
            self._where[cid] = base + j
        self.chunk_id.extend(chunk_ids)
        self.user_id.extend(user_ids)
        self.document_id.extend(doc_ids)
        self.modality.extend(modalities)
        self.meta.extend(metas)
        self._alive.extend([True] * len(chunk_ids))
        # duplicates inside one batch: only the last one stays alive (delete-all-then-add would keep both in
        # LanceDB; a chunk id is a primary key everywhere else in the reference, so we keep it unique)
        for j, cid in enumerate(chunk_ids):
            if self._where[cid] != base + j:
                self._alive[base + j] = False
        self._blocks.append(np.ascontiguousarray(emb, dtype=np.float32))
        self._dirty = True

    def upsert(self, rows: Iterable[VectorRow]) -> List[str]:
        rows = list(rows)
        if not rows:
            return []
        emb = [_unit_f32(r.embedding) for r in rows]
        dims = {e.shape[0] for e in emb}
        if len(dims) != 1:
            raise ValueError(f"{self.name}: embeddings of different lengths in one upsert: {sorted(dims)}")
        self._append([r.chunk_id for r in rows], [str(r.user_id) for r in rows], [r.document_id for r in rows],
                     [r.modality for r in rows], [json.dumps(r.meta or {}) for r in rows], np.stack(emb))
        return sorted({str(r.user_id) for r in rows})

    def load_columns(self, chunk_ids, user_ids, doc_ids, modalities, metas, emb: np.ndarray) -> None:
        """Bulk load of rows that are already normalised (what a LanceDB table holds)."""
        self._append(list(chunk_ids), [str(u) for u in user_ids], list(doc_ids), list(modalities), list(metas), emb)

    # -- resident copy ------------------------------------------------------------------------
    def _rebuild(self) -> None:
        alive = np.nonzero(np.asarray(self._alive, dtype=bool))[0]
        if alive.size == 0:
            self._resident, self._perm, self._seg_of = None, np.empty(0, np.int64), {}
            self._dirty = False
            return
        dims = {b.shape[1] for b in self._blocks}
        if len(dims) != 1:
            raise ValueError(f"{self.name}: rows of different embedding lengths {sorted(dims)} (the scan needs one dim)")
        mat = self._blocks[0] if len(self._blocks) == 1 else np.concatenate(self._blocks, axis=0)
        self._blocks = [mat]
        users = np.asarray(self.user_id, dtype=object)[alive]
        uniq, inv = np.unique(users.astype(str), return_inverse=True)
        order = np.argsort(inv, kind="stable")            # tenant-sorted, host order kept inside a tenant
        self._perm = alive[order]
        counts = np.bincount(inv, minlength=len(uniq))
        seg = np.zeros(len(uniq) + 1, dtype=np.int64)
        np.cumsum(counts, out=seg[1:])
        self._seg_of = {str(u): i for i, u in enumerate(uniq)}
        if self._resident is not None:
            self._resident.close()
        self._resident = ResidentIndex.from_f32(mat[self._perm], seg, dtype=self.dtype, device=self.device)
        self._dirty = False
        self.rebuilds += 1

    def resident(self) -> Optional[ResidentIndex]:
        if self._dirty:
            self._rebuild()
        return self._resident

    # -- reads --------------------------------------------------------------------------------
    def search(self, user_ids: Sequence[str], vectors: np.ndarray, top_k: int) -> List[List[Dict[str, Any]]]:
        limit = max(int(top_k), 1)
        if limit > N.MMR_MAX_K:
            raise N.NativeError(f"top_k {limit} > {N.MMR_MAX_K}: not supported by the resident-index kernels")
        res = self.resident()
        out: List[List[Dict[str, Any]]] = [[] for _ in user_ids]
        if res is None:
            return out
        segs = [self._seg_of.get(str(u), -1) for u in user_ids]
        live = [i for i, s in enumerate(segs) if s >= 0]
        if not live:
            return out
        q = np.ascontiguousarray(vectors[live], dtype=np.float32)
        scores, rows = res.search_host(q, limit, [segs[i] for i in live])
        one = np.float32(1.0)
        for j, i in enumerate(live):
            hits = []
            for s, r in zip(scores[j], rows[j]):
                if r < 0:
                    break
                h = int(self._perm[r])
                distance = float(one - s)                 # Lance returns the f32 cosine distance
                hits.append({
                    "chunk_id": self.chunk_id[h],
                    "score": 1.0 - distance,              # _format_results (:130-131)
                    "meta": json.loads(self.meta[h] or "{}"),
                })
            out[i] = hits
        return out


class B200Store:
    """Drop-in for `app.ml.retrieve._LANCEDB_STORE` / `app.ml.index_build._LANCEDB_STORE`."""

    def __init__(self, db_path: Optional[str] = None, device: Any = "cuda:0", dtype: str = "bf16") -> None:
        N.lib()  # fail now, loudly, if the CUDA library is missing
        if not torch.cuda.is_available():
            raise N.NativeError("no CUDA device: B200Store has no CPU fallback")
        self._device = torch.device(device)
        self._text_table = _Collection("text_collection", self._device, dtype)
        self._image_table = _Collection("image_collection", self._device, dtype)
        self._versions = VersionFile(os.path.join(db_path, "index_versions.json") if db_path else None)

    # writes (lancedb_store.py:87-101) + version bump (index_build.py:102,148)
    def upsert_text_vectors(self, rows: Iterable[VectorRow]) -> None:
        for user in self._text_table.upsert(rows):
            self._versions.bump(user)

    def upsert_image_vectors(self, rows: Iterable[VectorRow]) -> None:
        for user in self._image_table.upsert(rows):
            self._versions.bump(user)

    def get_index_version(self, user_id: str) -> int:
        return self._versions.get(user_id)

    def load_arrow(self, collection: str, table) -> None:
        """Bulk-load a pyarrow Table with the reference schema (chunk_id, user_id, document_id, modality,
        embedding: list<float32>, meta) -- what `lancedb.Table.to_arrow()` yields."""
        coll = {"text_collection": self._text_table, "image_collection": self._image_table}[collection]
        n = table.num_rows
        if n == 0:
            return
        emb = table.column("embedding").combine_chunks()
        offsets = emb.offsets.to_numpy()
        widths = np.diff(offsets)
        if widths.size and (widths != widths[0]).any():
            raise ValueError(f"{collection}: variable-length embeddings are not supported by the resident scan")
        flat = emb.values.to_numpy(zero_copy_only=False)[offsets[0]:offsets[-1]]
        mat = np.ascontiguousarray(flat, dtype=np.float32).reshape(n, int(widths[0]))
        cols = {c: table.column(c).to_pylist() for c in ("chunk_id", "user_id", "document_id", "modality", "meta")}
        coll.load_columns(cols["chunk_id"], cols["user_id"], cols["document_id"], cols["modality"], cols["meta"], mat)
        for user in sorted(set(map(str, cols["user_id"]))):
            self._versions.bump(user)

    # reads (lancedb_store.py:103-123)
    def search_text(self, user_id: str, query_vec: Sequence[float], top_k: int) -> List[Dict[str, Any]]:
        q = np.asarray(query_vec, dtype=np.float32)[None, :]
        return self._text_table.search([user_id], q, top_k)[0]

    def search_image(self, user_id: str, query_vec: Sequence[float], top_k: int) -> List[Dict[str, Any]]:
        q = np.asarray(query_vec, dtype=np.float32)[None, :]
        return self._image_table.search([user_id], q, top_k)[0]

    # micro-batched requests: one launch for B (tenant, query) pairs
    def search_text_batch(self, user_ids: Sequence[str], query_vecs, top_k: int) -> List[List[Dict[str, Any]]]:
        return self._text_table.search(list(user_ids), np.asarray(query_vecs, dtype=np.float32), top_k)

    def search_image_batch(self, user_ids: Sequence[str], query_vecs, top_k: int) -> List[List[Dict[str, Any]]]:
        return self._image_table.search(list(user_ids), np.asarray(query_vecs, dtype=np.float32), top_k)


def make_arrow_table(chunk_ids, user_ids, document_ids, modalities, embeddings: np.ndarray, metas):
    """A pyarrow Table with exactly the reference schema (lancedb_store.py:33-44); fixture / export helper."""
    import pyarrow as pa

    emb = np.ascontiguousarray(embeddings, dtype=np.float32)
    n, d = emb.shape
    offsets = pa.array(np.arange(0, (n + 1) * d, d, dtype=np.int32))
    lst = pa.ListArray.from_arrays(offsets, pa.array(emb.reshape(-1), type=pa.float32()))
    schema = pa.schema([
        pa.field("chunk_id", pa.string()), pa.field("user_id", pa.string()), pa.field("document_id", pa.string()),
        pa.field("modality", pa.string()), pa.field("embedding", pa.list_(pa.float32())),
        pa.field("meta", pa.string(), nullable=True),
    ])
    return pa.Table.from_arrays(
        [pa.array(list(chunk_ids), pa.string()), pa.array(list(user_ids), pa.string()),
         pa.array(list(document_ids), pa.string()), pa.array(list(modalities), pa.string()), lst,
         pa.array(list(metas), pa.string())], schema=schema)
