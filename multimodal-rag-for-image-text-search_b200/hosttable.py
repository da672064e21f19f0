"""HostTable -- columnar host master copy of one collection, sized for the BASELINE tables (10M-50M rows).

This is the host half of the resident-index store: what the reference keeps inside a LanceDB table
(app/storage/lancedb_store.py:33-44, six columns) is held here as Arrow / numpy columns, never as Python lists or
per-row dicts:

  * embeddings    one f32 [n_i, D] numpy view per appended block -- for an Arrow table that is a ZERO-COPY view of the
                  `embedding` column's value buffer (memory-mapped when the table came from an IPC file); blocks are
                  never concatenated, so an upsert costs O(new rows)
  * string columns (chunk_id, user_id, document_id, modality, meta)  one pyarrow StringArray per block
  * per-row state `alive` (bool), `tenant` (int32 id of the row's user_id), `hash` (u64 of chunk_id) as growable
                  numpy arrays
  * delete-by-chunk_id (lancedb_store.py:91-92) through a two-level sorted (hash, row) index searched with
                  np.searchsorted -- strings are compared only on hash hits

Pure host code (numpy + pyarrow + one host-only helper of the C library for string hashing): usable and tested without
a GPU.  The GPU half (`store._Collection`) decides where each alive row lives in HBM.
"""
from __future__ import annotations

import json
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

STRING_COLS = ("chunk_id", "user_id", "document_id", "modality", "meta")
_L1_MAX = 262_144  # second-level index is merged into the first when it grows past this


def _pa():
    import pyarrow as pa
    return pa


def hash_strings(arr) -> np.ndarray:
    """u64 hash per element of a pyarrow StringArray (nulls hash as "")."""
    from . import _native as N

    pa = _pa()
    n = len(arr)
    out = np.empty(n, dtype=np.uint64)
    if n == 0:
        return out
    if arr.type != pa.string():
        arr = arr.cast(pa.string())
    if arr.null_count:
        arr = arr.fill_null("")
    bufs = arr.buffers()  # [validity, offsets(int32), data]
    offsets = np.frombuffer(bufs[1], dtype=np.int32, count=n + 1, offset=arr.offset * 4)
    data_ptr = bufs[2].address if bufs[2] is not None else 0
    N.check(N.lib().mmr_hash_strings(data_ptr, offsets.ctypes.data, n, out.ctypes.data))
    return out


def _grow(arr: np.ndarray, need: int) -> np.ndarray:
    if need <= arr.shape[0]:
        return arr
    out = np.empty(max(need, int(arr.shape[0] * 1.5) + 1024), dtype=arr.dtype)
    out[: arr.shape[0]] = arr
    return out


class _Block:
    __slots__ = ("start", "n", "emb", "cols")

    def __init__(self, start: int, n: int, emb: np.ndarray, cols: Dict[str, Any]) -> None:
        self.start, self.n, self.emb, self.cols = start, n, emb, cols


def embedding_matrix(table) -> np.ndarray:
    """`embedding: list<float32>` column -> f32 [n, D] numpy array, zero-copy when the column is one chunk."""
    pa = _pa()
    col = table.column("embedding")
    emb = col.chunk(0) if col.num_chunks == 1 else col.combine_chunks()
    n = len(emb)
    if n == 0:
        return np.zeros((0, 0), dtype=np.float32)
    if pa.types.is_fixed_size_list(emb.type):
        d = emb.type.list_size
        flat = emb.values.to_numpy(zero_copy_only=False)[emb.offset * d:(emb.offset + n) * d]
    else:
        offsets = emb.offsets.to_numpy()
        widths = np.diff(offsets)
        if widths.size and (widths != widths[0]).any():
            raise ValueError("variable-length embeddings are not supported by the resident scan")
        d = int(widths[0])
        flat = emb.values.to_numpy(zero_copy_only=False)[offsets[0]:offsets[-1]]
    if flat.dtype != np.float32:
        flat = flat.astype(np.float32)
    return flat.reshape(n, d)


class HostTable:
    def __init__(self, name: str) -> None:
        self.name = name
        self.dim: Optional[int] = None
        self.blocks: List[_Block] = []
        self._starts = np.zeros(0, dtype=np.int64)
        self._starts_list: List[int] = []
        self.n_total = 0
        self.n_alive = 0
        self.alive = np.zeros(0, dtype=bool)
        self.tenant = np.zeros(0, dtype=np.int32)
        self.hash = np.zeros(0, dtype=np.uint64)
        self.tenants: List[str] = []
        self._tenant_id: Dict[str, int] = {}
        # sorted (hash, row) index: level 0 = big, rebuilt rarely; level 1 = recent appends
        self._l0_hash = np.zeros(0, dtype=np.uint64)
        self._l0_row = np.zeros(0, dtype=np.int64)
        self._l1_hash = np.zeros(0, dtype=np.uint64)
        self._l1_row = np.zeros(0, dtype=np.int64)
        self._indexed_upto = 0   # rows below this are in l0 / l1; the rest are indexed lazily on the next lookup

    def __len__(self) -> int:
        return self.n_alive

    # ------------------------------------------------------------------------------------------ lookups
    def _block_of(self, row: int) -> Tuple[_Block, int]:
        b = int(np.searchsorted(self._starts, row, side="right")) - 1
        blk = self.blocks[b]
        return blk, row - blk.start

    def value_at(self, col: str, row: int):
        blk, j = self._block_of(int(row))
        return blk.cols[col][j].as_py()

    def chunk_id_at(self, row: int) -> str:
        return self.value_at("chunk_id", row)

    def meta_at(self, row: int) -> Optional[str]:
        return self.value_at("meta", row)

    def _raw(self, blk: _Block, col: str):
        """(int32 offsets, memoryview of the utf-8 bytes) of one string column of one block, cached; None when the
        column has nulls (then values are read through Arrow scalars)."""
        cache = blk.cols.setdefault("_raw", {})
        hit = cache.get(col, 0)
        if hit == 0:
            arr = blk.cols[col]
            if arr.null_count or len(arr) == 0:
                hit = None
            else:
                bufs = arr.buffers()
                # offsets as a memoryview of C ints: indexing yields Python ints directly (a numpy scalar per offset made
                # the ten lookups of a result list cost ~20 us)
                off = memoryview(bufs[1])[arr.offset * 4:(arr.offset + len(arr) + 1) * 4].cast("i")
                hit = (off, memoryview(bufs[2]) if bufs[2] is not None else memoryview(b""))
            cache[col] = hit
        return hit

    def values_at(self, col: str, rows: Sequence[int]) -> list:
        """Column values of several host rows, read straight out of the Arrow buffers (offset pair + one utf-8 decode per
        value): the metadata side of a hit costs about a microsecond, with no Arrow scalar / compute-kernel / numpy
        small-array round trip (a result list is ~10 rows: plain Python beats vectorisation here)."""
        import bisect

        out = []
        starts = self._starts_list
        single = len(self.blocks) == 1
        last_b, blk, raw, start = -1, None, None, 0
        for r in rows:
            r = int(r)
            b = 0 if single else bisect.bisect_right(starts, r) - 1
            if b != last_b:
                blk = self.blocks[b]
                raw, start, last_b = self._raw(blk, col), blk.start, b
            j = r - start
            if raw is None:
                out.append(blk.cols[col][j].as_py())
            else:
                off, data = raw
                out.append(str(data[off[j]:off[j + 1]], "utf-8"))
        return out

    def hits_at(self, rows: Sequence[int]) -> Tuple[list, list]:
        """(chunk ids, parsed meta dicts) of a result list's host rows in ONE pass over the Arrow buffers; the empty meta
        ("{}" / "" / null -- what ingest writes for most chunks) is recognised on its raw bytes without decoding or parsing
        (reference _format_results, lancedb_store.py:125-139: json.loads(meta) if meta else {})."""
        import bisect
        import json

        ids, metas = [], []
        starts = self._starts_list
        single = len(self.blocks) == 1
        last_b, blk, raw_id, raw_meta, start = -1, None, None, None, 0
        for r in rows:
            r = int(r)
            b = 0 if single else bisect.bisect_right(starts, r) - 1
            if b != last_b:
                blk = self.blocks[b]
                raw_id, raw_meta, start, last_b = self._raw(blk, "chunk_id"), self._raw(blk, "meta"), blk.start, b
            j = r - start
            if raw_id is None:
                ids.append(blk.cols["chunk_id"][j].as_py())
            else:
                off, data = raw_id
                ids.append(str(data[off[j]:off[j + 1]], "utf-8"))
            if raw_meta is None:
                m = blk.cols["meta"][j].as_py()
                metas.append({} if m in (None, "", "{}") else json.loads(m))
            else:
                off, data = raw_meta
                a, e = off[j], off[j + 1]
                if e == a or (e - a == 2 and data[a] == 123 and data[a + 1] == 125):   # "" or "{}" (byte tests: comparing a
                    metas.append({})                                                   # memoryview with bytes is slow)
                else:
                    metas.append(json.loads(str(data[a:e], "utf-8")))
        return ids, metas

    def gather(self, rows: np.ndarray) -> np.ndarray:
        """f32 embeddings of the given host rows (any order), O(len(rows))."""
        rows = np.asarray(rows, dtype=np.int64)
        out = np.empty((rows.shape[0], self.dim or 0), dtype=np.float32)
        if rows.size == 0:
            return out
        which = np.searchsorted(self._starts, rows, side="right") - 1
        for b in np.unique(which):
            sel = np.nonzero(which == b)[0]
            out[sel] = self.blocks[int(b)].emb[rows[sel] - self.blocks[int(b)].start]
        return out

    def tenant_id(self, user_id: str) -> Optional[int]:
        return self._tenant_id.get(str(user_id))

    def _ensure_index(self) -> None:
        """Bring the (hash, row) index up to date with every appended row."""
        if self._indexed_upto == self.n_total:
            return
        lo, hi = self._indexed_upto, self.n_total
        new_h = self.hash[lo:hi]
        new_r = np.arange(lo, hi, dtype=np.int64)
        if (hi - lo) + self._l1_hash.shape[0] > _L1_MAX or self._l0_hash.shape[0] == 0:
            h = np.concatenate([self._l0_hash, self._l1_hash, new_h])
            r = np.concatenate([self._l0_row, self._l1_row, new_r])
            order = np.argsort(h, kind="stable")
            self._l0_hash, self._l0_row = h[order], r[order]
            self._l1_hash, self._l1_row = np.zeros(0, np.uint64), np.zeros(0, np.int64)
        else:
            h = np.concatenate([self._l1_hash, new_h])
            r = np.concatenate([self._l1_row, new_r])
            order = np.argsort(h, kind="stable")
            self._l1_hash, self._l1_row = h[order], r[order]
        self._indexed_upto = hi

    def find_alive(self, ids, hashes: Optional[np.ndarray] = None) -> np.ndarray:
        """Host row of the ALIVE row with each chunk_id, -1 when there is none."""
        pa = _pa()
        arr = ids if isinstance(ids, pa.Array) else pa.array(list(ids), pa.string())
        n = len(arr)
        out = np.full(n, -1, dtype=np.int64)
        if n == 0 or self.n_total == 0:
            return out
        if hashes is None:
            hashes = hash_strings(arr)
        self._ensure_index()
        for lvl_h, lvl_r in ((self._l0_hash, self._l0_row), (self._l1_hash, self._l1_row)):
            if lvl_h.shape[0] == 0:
                continue
            lo = np.searchsorted(lvl_h, hashes, side="left")
            hi = np.searchsorted(lvl_h, hashes, side="right")
            for i in np.nonzero(hi > lo)[0]:
                want = None
                for j in range(int(lo[i]), int(hi[i])):
                    row = int(lvl_r[j])
                    if not self.alive[row]:
                        continue
                    if want is None:
                        want = arr[int(i)].as_py()
                    if self.chunk_id_at(row) == want:
                        out[i] = row
                        break
        return out

    # ------------------------------------------------------------------------------------------ writes
    def append_table(self, table) -> Tuple[int, int, np.ndarray, List[str]]:
        """Append the rows of a pyarrow Table with the reference schema (embeddings already unit-norm).

        Upsert semantics (lancedb_store.py:91-93): an alive row with the same chunk_id is deleted first; when a chunk_id
        occurs several times inside the batch the last occurrence wins.
        Returns (first new host row, number of rows appended, host rows that were alive and are now dead, users touched).
        """
        pa = _pa()
        import pyarrow.compute as pc

        n = table.num_rows
        if n == 0:
            return self.n_total, 0, np.zeros(0, np.int64), []
        emb = embedding_matrix(table)
        if self.dim is None:
            self.dim = int(emb.shape[1])
        elif emb.shape[1] != self.dim:
            raise ValueError(f"{self.name}: embedding length {emb.shape[1]} != {self.dim} already stored")
        cols = {}
        for c in STRING_COLS:
            col = table.column(c)
            a = col.chunk(0) if col.num_chunks == 1 else col.combine_chunks()
            cols[c] = a if a.type == pa.string() else a.cast(pa.string())
        base = self.n_total
        hashes = hash_strings(cols["chunk_id"])
        # rows this batch replaces (looked up BEFORE the batch becomes visible to the index)
        old = self.find_alive(cols["chunk_id"], hashes) if self.n_total else np.full(n, -1, np.int64)
        killed = np.unique(old[old >= 0])
        # duplicates inside the batch: all but the last occurrence are dead on arrival
        alive_new = np.ones(n, dtype=bool)
        order = np.argsort(hashes, kind="stable")
        hs = hashes[order]
        dup = np.nonzero(hs[1:] == hs[:-1])[0]
        if dup.size:
            ids = cols["chunk_id"]
            for d in dup:                       # order is stable: order[d] < order[d+1] within an equal-hash run
                a, b = int(order[d]), int(order[d + 1])
                if ids[a].as_py() == ids[b].as_py():
                    alive_new[a] = False
                else:                           # true hash collision inside a run: compare against every later member
                    j = d + 1
                    while j + 1 < n and hs[j + 1] == hs[d]:
                        j += 1
                        if ids[a].as_py() == ids[int(order[j])].as_py():
                            alive_new[a] = False
                            break
        # tenants
        enc = pc.dictionary_encode(cols["user_id"].fill_null(""))
        names = enc.dictionary.to_pylist()
        lut = np.empty(max(len(names), 1), dtype=np.int32)
        for i, name in enumerate(names):
            tid = self._tenant_id.get(name)
            if tid is None:
                tid = len(self.tenants)
                self.tenants.append(name)
                self._tenant_id[name] = tid
            lut[i] = tid
        tenant_new = lut[enc.indices.to_numpy(zero_copy_only=False).astype(np.int64)]
        # commit
        self.alive = _grow(self.alive, base + n)
        self.tenant = _grow(self.tenant, base + n)
        self.hash = _grow(self.hash, base + n)
        self.alive[base:base + n] = alive_new
        self.tenant[base:base + n] = tenant_new
        self.hash[base:base + n] = hashes
        if killed.size:
            self.alive[killed] = False
        self.blocks.append(_Block(base, n, emb, cols))
        self._starts = np.append(self._starts, base)
        self._starts_list.append(int(base))
        self.n_total = base + n
        self.n_alive += int(alive_new.sum()) - int(killed.size)
        return base, n, killed, sorted(set(names))

    def alive_rows(self) -> np.ndarray:
        return np.nonzero(self.alive[: self.n_total])[0]

    def to_arrow(self):
        """Alive rows, host order, as one pyarrow Table with the reference schema (persist / compaction)."""
        pa = _pa()
        parts = []
        for blk in self.blocks:
            mask = self.alive[blk.start:blk.start + blk.n]
            if not mask.any():
                continue
            sel = np.nonzero(mask)[0]
            emb = blk.emb if sel.size == blk.n else blk.emb[sel]
            take = None if sel.size == blk.n else pa.array(sel)
            cols = {c: (blk.cols[c] if take is None else blk.cols[c].take(take)) for c in STRING_COLS}
            parts.append(make_arrow_table(cols["chunk_id"], cols["user_id"], cols["document_id"], cols["modality"], emb,
                                          cols["meta"]))
        if not parts:
            return None
        return pa.concat_tables(parts) if len(parts) > 1 else parts[0]


def make_arrow_table(chunk_ids, user_ids, document_ids, modalities, embeddings: np.ndarray, metas):
    """A pyarrow Table with exactly the reference schema (lancedb_store.py:33-44); fixture / export helper."""
    pa = _pa()
    emb = np.ascontiguousarray(embeddings, dtype=np.float32)
    n, d = emb.shape
    if n * d < 2 ** 31 - 1:
        offsets = pa.array(np.arange(0, (n + 1) * d, d, dtype=np.int32))
        lst = pa.ListArray.from_arrays(offsets, pa.array(emb.reshape(-1), type=pa.float32()))
        emb_type = pa.list_(pa.float32())
    else:  # list<float32> offsets are int32: beyond 2^31 values the column must be large_list
        offsets = pa.array(np.arange(0, (n + 1) * d, d, dtype=np.int64))
        lst = pa.LargeListArray.from_arrays(offsets, pa.array(emb.reshape(-1), type=pa.float32()))
        emb_type = pa.large_list(pa.float32())

    def strings(v):
        return v if isinstance(v, pa.Array) else pa.array(list(v), pa.string())

    schema = pa.schema([
        pa.field("chunk_id", pa.string()), pa.field("user_id", pa.string()), pa.field("document_id", pa.string()),
        pa.field("modality", pa.string()), pa.field("embedding", emb_type),
        pa.field("meta", pa.string(), nullable=True),
    ])
    return pa.Table.from_arrays([strings(chunk_ids), strings(user_ids), strings(document_ids), strings(modalities), lst,
                                 strings(metas)], schema=schema)


def rows_to_arrow(rows: Sequence[Any], normalize) -> Any:
    """VectorRow payloads -> Arrow table, embeddings normalised like LanceDBStore._prepare_rows (:71-85)."""
    emb = [normalize(r.embedding) for r in rows]
    dims = {e.shape[0] for e in emb}
    if len(dims) != 1:
        raise ValueError(f"embeddings of different lengths in one upsert: {sorted(dims)}")
    return make_arrow_table([r.chunk_id for r in rows], [str(r.user_id) for r in rows], [r.document_id for r in rows],
                            [r.modality for r in rows], np.stack(emb), [json.dumps(r.meta or {}) for r in rows])
