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
    """Host master copy (columns in insertion order) + the resident copy on the GPU.

    Resident layout = [base: tenant-sorted rows of the last compaction][delta segments appended since].  An upsert
    tombstones the replaced rows (their resident vectors are overwritten with NaN, which the kernels never
    return) and appends the new rows grouped by tenant, so a tenant owns a short list of row ranges and every
    search scans exactly those (mmr_search_ranges).  A compaction (full tenant-sorted rebuild) runs when the
    buffer is full, tombstones exceed a quarter of it, or a tenant has collected more than MAX_RANGES ranges.
    This is the incremental refresh path of SURVEY 8f: an upsert no longer re-uploads the table.
    """

    MAX_RANGES = 8
    MAX_TOMB_FRACTION = 0.25

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
        self._block_start: List[int] = []
        self._alive: List[bool] = []
        self._where: Dict[str, int] = {}        # chunk_id -> host row (alive rows only)
        # resident state
        self._resident: Optional[ResidentIndex] = None
        self._buf: Optional[torch.Tensor] = None   # capacity buffer [cap, D]
        self._n_res = 0
        self._perm = np.empty(0, np.int64)         # resident row -> host row
        self._res_of: Dict[int, int] = {}          # host row -> resident row
        self._ranges: Dict[str, List[List[int]]] = {}
        self._tomb = 0
        self._pending_rows: List[int] = []         # host rows not yet resident
        self._pending_tomb: List[int] = []         # resident rows to overwrite with NaN
        self._force_rebuild = True
        self._seg_key = None
        self.rebuilds = 0
        self.appends = 0

    # -- writes -------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self._where)

    def _kill(self, host_row: int) -> None:
        self._alive[host_row] = False           # table.delete("chunk_id == '...'") (lancedb_store.py:91-92)
        res = self._res_of.pop(host_row, None)
        if res is not None:
            self._pending_tomb.append(res)

    def _append(self, chunk_ids, user_ids, doc_ids, modalities, metas, emb: np.ndarray) -> None:
        base = len(self.chunk_id)
        emb = np.ascontiguousarray(emb, dtype=np.float32)
        if self._blocks and emb.shape[1] != self._blocks[0].shape[1]:
            raise ValueError(f"{self.name}: embedding length {emb.shape[1]} != {self._blocks[0].shape[1]} already stored")
        self.chunk_id.extend(chunk_ids)
        self.user_id.extend(user_ids)
        self.document_id.extend(doc_ids)
        self.modality.extend(modalities)
        self.meta.extend(metas)
        self._alive.extend([True] * len(chunk_ids))
        for j, cid in enumerate(chunk_ids):
            old = self._where.get(cid)
            if old is not None:
                # an id is a primary key everywhere else in the reference: the newest row wins (also inside a batch)
                self._kill(old)
            self._where[cid] = base + j
        self._block_start.append(base)
        self._blocks.append(emb)
        self._pending_rows.extend(base + j for j in range(len(chunk_ids)) if self._alive[base + j])

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

    def _host_rows(self, idx: np.ndarray) -> np.ndarray:
        """f32 embeddings of the given host rows."""
        if len(self._blocks) > 1:
            self._blocks = [np.concatenate(self._blocks, axis=0)]
            self._block_start = [0]
        return self._blocks[0][idx]

    # -- resident copy ------------------------------------------------------------------------
    def _rebuild(self) -> None:
        """Compaction: tenant-sorted base, no deltas, no tombstones."""
        alive = np.nonzero(np.asarray(self._alive, dtype=bool))[0]
        self._pending_rows, self._pending_tomb, self._tomb = [], [], 0
        self._force_rebuild = False
        if self._resident is not None:
            self._resident.close()
            self._resident = None
        if alive.size == 0:
            self._buf, self._n_res, self._perm, self._res_of, self._ranges = None, 0, np.empty(0, np.int64), {}, {}
            return
        users = np.asarray(self.user_id, dtype=object)[alive].astype(str)
        uniq, inv = np.unique(users, return_inverse=True)
        order = np.argsort(inv, kind="stable")            # tenant-sorted, host order kept inside a tenant
        self._perm = alive[order]
        counts = np.bincount(inv, minlength=len(uniq))
        seg = np.zeros(len(uniq) + 1, dtype=np.int64)
        np.cumsum(counts, out=seg[1:])
        n = int(alive.size)
        cap = max(n + 4096, int(n * 1.25))
        base = ResidentIndex.from_f32(self._host_rows(self._perm), None, dtype=self.dtype, device=self.device)
        self._buf = torch.empty((cap, base.dim), dtype=base.rows.dtype, device=self.device)
        self._buf[:n].copy_(base.rows)
        base.close()
        self._n_res = n
        self._resident = ResidentIndex(self._buf[:n])
        self._resident.update(self._buf, n)
        self._seg_key = None
        self._res_of = {int(h): i for i, h in enumerate(self._perm)}
        self._ranges = {str(u): [[int(seg[i]), int(seg[i + 1])]] for i, u in enumerate(uniq)}
        self.rebuilds += 1

    def _apply_deltas(self) -> None:
        """Tombstone replaced rows and append the pending rows as per-tenant delta segments."""
        if self._pending_tomb:
            idx = torch.as_tensor(self._pending_tomb, dtype=torch.int64, device=self.device)
            self._buf.index_fill_(0, idx, float("nan"))
            self._tomb += len(self._pending_tomb)
            self._pending_tomb = []
        rows = [h for h in self._pending_rows if self._alive[h]]
        self._pending_rows = []
        if rows:
            rows = np.asarray(rows, dtype=np.int64)
            users = np.asarray([self.user_id[h] for h in rows], dtype=object).astype(str)
            order = np.argsort(users, kind="stable")       # group the batch by tenant, arrival order inside
            rows, users = rows[order], users[order]
            m = len(rows)
            lo = self._n_res
            delta = ResidentIndex.from_f32(self._host_rows(rows), None, dtype=self.dtype, device=self.device)
            self._buf[lo:lo + m].copy_(delta.rows)
            delta.close()
            self._perm = np.concatenate([self._perm, rows])
            for j, h in enumerate(rows):
                self._res_of[int(h)] = lo + j
            start = 0
            while start < m:
                end = start
                while end < m and users[end] == users[start]:
                    end += 1
                rl = self._ranges.setdefault(str(users[start]), [])
                if rl and rl[-1][1] == lo + start:
                    rl[-1][1] = lo + end                      # contiguous with the tenant's last range
                else:
                    rl.append([lo + start, lo + end])
                start = end
            self._n_res = lo + m
            self.appends += 1
        self._resident.update(self._buf, self._n_res)
        self._seg_key = None

    def resident(self) -> Optional[ResidentIndex]:
        if not self._force_rebuild and not self._pending_rows and not self._pending_tomb:
            return self._resident
        n_new = sum(1 for h in self._pending_rows if self._alive[h])
        need_rebuild = (
            self._force_rebuild or self._resident is None or self._buf is None
            or self._n_res + n_new > self._buf.shape[0]
            or (self._tomb + len(self._pending_tomb)) > self.MAX_TOMB_FRACTION * max(self._n_res, 1)
            or any(len(r) >= self.MAX_RANGES for r in self._ranges.values())
        )
        if need_rebuild:
            self._rebuild()
        else:
            self._apply_deltas()
        return self._resident

    # -- reads --------------------------------------------------------------------------------
    def search_device(self, user_ids: Sequence[str], vectors: np.ndarray, top_k: int):
        """Batched search that leaves the results on the device: (scores [B,k] f32, rows [B,k] i64) with -1 rows for
        unknown tenants, or None when the collection is empty."""
        limit = max(int(top_k), 1)
        if limit > N.MMR_MAX_K:
            raise N.NativeError(f"top_k {limit} > {N.MMR_MAX_K}: not supported by the resident-index kernels")
        res = self.resident()
        if res is None:
            return None
        ranges = [self._ranges.get(str(u)) or [] for u in user_ids]
        q = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32)).to(self.device)
        return res.search_ranges(q, limit, ranges)

    def row_identity(self, resident_row: int) -> int:
        return int(self._perm[resident_row])

    def search(self, user_ids: Sequence[str], vectors: np.ndarray, top_k: int) -> List[List[Dict[str, Any]]]:
        limit = max(int(top_k), 1)
        if limit > N.MMR_MAX_K:
            raise N.NativeError(f"top_k {limit} > {N.MMR_MAX_K}: not supported by the resident-index kernels")
        res = self.resident()
        out: List[List[Dict[str, Any]]] = [[] for _ in user_ids]
        if res is None:
            return out
        ranges = [self._ranges.get(str(u)) for u in user_ids]
        live = [i for i, r in enumerate(ranges) if r]
        if not live:
            return out
        q = np.ascontiguousarray(vectors[live], dtype=np.float32)
        if all(len(ranges[i]) == 1 and ranges[i] == ranges[live[0]] for i in live):
            # one shared range: host-buffer C call (H2D + scan + D2H inside the library)
            lo, hi = ranges[live[0]][0]
            if self._seg_key != (lo, hi, self._n_res):          # re-point the one-segment table only when it changes
                res.update(self._buf, self._n_res, seg_offsets=[lo, hi])
                self._seg_key = (lo, hi, self._n_res)
            scores, rows = res.search_host(q, limit, [0] * len(live))
        else:
            s_dev, r_dev = res.search_ranges(torch.from_numpy(q).to(self.device), limit, [ranges[i] for i in live])
            scores, rows = s_dev.cpu().numpy(), r_dev.cpu().numpy()
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
        self._db_path = db_path
        self._text_table = _Collection("text_collection", self._device, dtype)
        self._image_table = _Collection("image_collection", self._device, dtype)
        self._versions = VersionFile(os.path.join(db_path, "index_versions.json") if db_path else None)
        if db_path:
            # durable state next to index_versions.json: one Arrow IPC file per collection with the reference schema
            # (what LanceDB keeps as .lance fragments); a restarted process reloads them into HBM
            for coll in (self._text_table, self._image_table):
                path = os.path.join(db_path, coll.name + ".arrow")
                if os.path.exists(path):
                    self._load_ipc(coll, path)

    @staticmethod
    def _load_ipc(coll: "_Collection", path: str) -> None:
        import pyarrow as pa
        import pyarrow.ipc as ipc

        with pa.memory_map(path, "r") as src:
            table = ipc.open_file(src).read_all()
        if table.num_rows:
            B200Store._load_table(coll, table)

    @staticmethod
    def _load_table(coll: "_Collection", table) -> List[str]:
        emb = table.column("embedding").combine_chunks()
        offsets = emb.offsets.to_numpy()
        widths = np.diff(offsets)
        if widths.size and (widths != widths[0]).any():
            raise ValueError(f"{coll.name}: variable-length embeddings are not supported by the resident scan")
        flat = emb.values.to_numpy(zero_copy_only=False)[offsets[0]:offsets[-1]]
        mat = np.ascontiguousarray(flat, dtype=np.float32).reshape(table.num_rows, int(widths[0]))
        cols = {c: table.column(c).to_pylist() for c in ("chunk_id", "user_id", "document_id", "modality", "meta")}
        coll.load_columns(cols["chunk_id"], cols["user_id"], cols["document_id"], cols["modality"], cols["meta"], mat)
        return sorted(set(map(str, cols["user_id"])))

    def load_arrow_ipc(self, collection: str, path: str) -> None:
        """Bulk-load an Arrow IPC file written with the reference schema (see make_arrow_table / persist)."""
        import pyarrow as pa
        import pyarrow.ipc as ipc

        with pa.memory_map(path, "r") as src:
            self.load_arrow(collection, ipc.open_file(src).read_all())

    def persist(self) -> None:
        """Write both collections (live rows, host order) as Arrow IPC files under db_path."""
        import pyarrow as pa
        import pyarrow.ipc as ipc

        if not self._db_path:
            raise ValueError("B200Store was created without a db_path")
        os.makedirs(self._db_path, exist_ok=True)
        for coll in (self._text_table, self._image_table):
            alive = np.nonzero(np.asarray(coll._alive, dtype=bool))[0]
            if alive.size == 0:
                continue
            table = make_arrow_table([coll.chunk_id[i] for i in alive], [coll.user_id[i] for i in alive],
                                     [coll.document_id[i] for i in alive], [coll.modality[i] for i in alive],
                                     coll._host_rows(alive), [coll.meta[i] for i in alive])
            tmp = os.path.join(self._db_path, coll.name + ".arrow.tmp")
            with pa.OSFile(tmp, "wb") as sink, ipc.new_file(sink, table.schema) as writer:
                writer.write_table(table)
            os.replace(tmp, os.path.join(self._db_path, coll.name + ".arrow"))

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
        if table.num_rows == 0:
            return
        for user in self._load_table(coll, table):
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

    def fused_search_batch(self, user_ids: Sequence[str], text_vecs, image_vecs, top_k_text: int, top_k_image: int,
                           final_n: int, tau: float):
        """Rerank-off request path kept on the device: text scan + image scan + z-score fusion + FINAL_N cut +
        CONFIDENCE_TAU gate (kernels K1/K2 + K5).  Returns per request ([{"chunk_id", "modality", "score",
        "combined_score"}, ...] best first, low_confidence flag) -- what retrieve() then _confidence_low() give when
        every hit survives the metadata join (reference app/ml/retrieve.py:103-117, app/ml/generate.py:56-60)."""
        return _fused_batch(self, list(user_ids), text_vecs, image_vecs, top_k_text, top_k_image, final_n, tau)


def _fused_batch(store: "B200Store", user_ids, text_vecs, image_vecs, kt: int, ki: int, final_n: int, tau: float):
    """scan(text) + scan(image) + K5 fusion/gate for a batch of requests, results fetched with one small D2H."""
    from .index import fuse

    t = store._text_table.search_device(user_ids, np.asarray(text_vecs, dtype=np.float32), kt)
    i = store._image_table.search_device(user_ids, np.asarray(image_vecs, dtype=np.float32), ki)
    if t is None and i is None:
        return [([], True) for _ in user_ids]
    out = fuse(t, i, final_n, tau)
    comb, score = out["combined"].cpu().numpy(), out["score"].cpu().numpy()
    rows, mod, low = out["rows"].cpu().numpy(), out["modality"].cpu().numpy(), out["low_conf"].cpu().numpy()
    results = []
    for b in range(len(user_ids)):
        items = []
        for o in range(final_n):
            if rows[b, o] < 0:
                break
            coll = store._text_table if mod[b, o] == 0 else store._image_table
            h = coll.row_identity(int(rows[b, o]))
            items.append({"chunk_id": coll.chunk_id[h], "modality": "text" if mod[b, o] == 0 else "image",
                          "score": float(score[b, o]), "combined_score": float(comb[b, o])})
        results.append((items, bool(low[b])))
    return results


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
