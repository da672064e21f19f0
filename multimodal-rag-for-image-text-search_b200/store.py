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
from array import array
import os
import threading
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .durable import DurableLog
from .hosttable import HostTable, make_arrow_table, rows_to_arrow
from .index import MultiIndex, ResidentIndex
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


def _grow_filled(arr: np.ndarray, need: int, fill) -> np.ndarray:
    if need <= arr.shape[0]:
        return arr
    out = np.full(max(need, int(arr.shape[0] * 1.5) + 1024), fill, dtype=arr.dtype)
    out[: arr.shape[0]] = arr
    return out


class _Collection:
    """Host master copy (`HostTable`: Arrow / numpy columns, O(new rows) per upsert) + the resident copy on the GPU.

    Resident layout = [base: tenant-sorted rows of the last compaction][delta segments appended since].  An upsert
    tombstones the replaced rows (their resident vectors are overwritten with NaN, which the kernels never
    return) and appends the new rows grouped by tenant, so a tenant owns a short list of row ranges and every
    search scans exactly those (mmr_search_ranges).  A compaction (full tenant-sorted rebuild) runs when the
    buffer is full, tombstones exceed a quarter of it, or a tenant has collected more than MAX_RANGES ranges; it
    streams every host block once through the scatter loader (insertion order on the host, tenant order in HBM).
    This is the incremental refresh path of SURVEY 8f: an upsert no longer re-uploads the table.

    With a `DurableLog` the collection is also the reader / writer of the shared on-disk state (durable.py): every
    upsert batch becomes a delta file before the version is bumped, and `refresh()` pulls in what other processes wrote.

    Thread safety: one re-entrant lock per collection serialises upserts, refreshes and searches (the resident
    workspace, the staging buffers of mmr_search_host and the pending-delta lists are per collection).
    """

    MAX_RANGES = 8
    MAX_TOMB_FRACTION = 0.25

    def __init__(self, name: str, device: torch.device, dtype: str, log: Optional[DurableLog] = None,
                 query_precision: str = "f32", devices: Optional[Sequence[torch.device]] = None) -> None:
        # devices = the GPUs of one box the collection is row-range-sharded over -- ONE process, one shard
        # and one launcher thread per device (index.MultiIndex).  Default: the single `device`.
        self.devices = [torch.device(d) for d in devices] if devices else [device]
        # "f32": what a request gets does not depend on which other requests shared its launch (the reference answers
        # every request alone); "auto" lets same-tenant batches of >= 3 run on the tensor cores with 16-bit queries
        self.query_precision = query_precision
        self.name = name
        self.device = device
        self.dtype = dtype
        self.host = HostTable(name)
        self._lock = threading.RLock()
        self._log = log
        self._gen: Optional[int] = None
        self._n_deltas = 0
        self._stamp = None
        self.rebuilds = 0
        self.appends = 0
        self.refreshes = 0
        self.last_load_gbs: Optional[float] = None   # L1 loader throughput of the last compaction (fp32 bytes / s)
        self._reset_resident()

    def _reset_resident(self) -> None:
        old = getattr(self, "_resident", None)
        if old is not None:
            old.close()
        for sh in getattr(self, "_shards", []) or []:
            sh.close()
        multi = getattr(self, "_multi", None)
        if multi is not None:
            multi.close()
        self._multi: Optional[MultiIndex] = None   # sharded form: G ResidentIndex shards behind one MultiIndex
        self._shards: List[ResidentIndex] = []
        self._bufs: List[torch.Tensor] = []
        self._bounds: List[int] = []               # global resident row where each shard starts (+ end of the base)
        self._resident: Optional[ResidentIndex] = None
        self._buf: Optional[torch.Tensor] = None   # capacity buffer [cap, D] (sharded: the LAST shard's buffer)
        self._n_res = 0
        self._perm = np.empty(0, np.int64)         # resident row -> host row
        self._res_of = np.empty(0, np.int64)       # host row -> resident row (-1 = not resident)
        self._ranges: Dict[str, List[List[int]]] = {}
        self._tomb = 0
        self._synced_upto = 0                      # host rows below this have been offered to the resident copy
        self._pending_tomb: List[int] = []         # resident rows to overwrite with NaN
        self._force_rebuild = True
        self._seg_key = None

    # -- writes -------------------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.host)

    def host_row_of(self, chunk_id: str) -> int:
        return int(self.host.find_alive([chunk_id])[0])

    def _host_rows(self, idx: np.ndarray) -> np.ndarray:
        """f32 embeddings of the given host rows."""
        return self.host.gather(np.asarray(idx, dtype=np.int64))

    def _ingest(self, table) -> List[str]:
        """Append an Arrow table (rows already unit-norm) to the host copy and queue the resident-side work."""
        _, n, killed, users = self.host.append_table(table)
        if n == 0:
            return []
        self._res_of = _grow_filled(self._res_of, self.host.n_total, -1)
        if killed.size:                           # table.delete("chunk_id == '...'") (lancedb_store.py:91-92)
            res = self._res_of[killed]
            self._pending_tomb.extend(int(r) for r in res[res >= 0])
            self._res_of[killed] = -1
        return users

    def upsert(self, rows: Iterable[VectorRow]) -> List[str]:
        rows = list(rows)
        if not rows:
            return []
        try:
            table = rows_to_arrow(rows, _unit_f32)
        except ValueError as exc:
            raise ValueError(f"{self.name}: {exc}") from None
        return self.add_table(table)

    def add_table(self, table) -> List[str]:
        """Upsert an Arrow table whose embeddings are already normalised (bulk load, or `upsert` after _prepare_rows).
        With a durable log the batch is on disk (and named by the manifest) before this returns."""
        if table.num_rows == 0:
            return []
        with self._lock:
            if self._log is None:
                return self._ingest(table)
            with self._log.locked():
                self.refresh()                    # another writer may have appended since we last looked
                users = self._ingest(table)
                self._log.append_delta(table)
                self._n_deltas += 1
                self._stamp = self._log.stamp()
            return users

    def refresh(self, force: bool = False, _depth: int = 0) -> List[str]:
        """Pull in what other processes made durable since the last look.  One os.stat when nothing changed."""
        if self._log is None:
            return []
        stamp = self._log.stamp()
        if stamp == self._stamp and not force and self._gen is not None:
            return []
        users: List[str] = []
        with self._lock:
            m = self._log.manifest()
            try:
                if m["generation"] != self._gen:
                    if self._gen is not None or self.host.n_total:
                        self.host = HostTable(self.name)
                        self._reset_resident()
                    if m["base"]:
                        users += self._ingest(self._log.read_table(m["base"]))
                    self._gen, self._n_deltas = m["generation"], 0
                for fname in m["deltas"][self._n_deltas:]:
                    users += self._ingest(self._log.read_table(fname))
                    self._n_deltas += 1
            except FileNotFoundError:
                if _depth >= 3:
                    raise
                self._gen = None                  # a compaction replaced the files under us: start over
                return self.refresh(force=True, _depth=_depth + 1)
            self._stamp = stamp
            self.refreshes += 1
        return sorted(set(users))

    def persist(self) -> None:
        """Durable compaction: all alive rows become the base of a new generation, delta files are dropped."""
        if self._log is None:
            raise ValueError("collection has no durable log")
        with self._lock, self._log.locked():
            self.refresh()
            m = self._log.write_base(self.host.to_arrow())
            self._gen, self._n_deltas = m["generation"], 0
            self._stamp = self._log.stamp()

    # -- resident copy ------------------------------------------------------------------------
    def _rebuild(self) -> None:
        """Compaction: tenant-sorted base, no deltas, no tombstones."""
        import time

        host = self.host
        alive = host.alive_rows()
        self._reset_resident()
        self._force_rebuild = False
        self._synced_upto = host.n_total
        self._res_of = np.full(host.n_total, -1, dtype=np.int64)
        n = int(alive.size)
        if n == 0:
            return
        tenant = host.tenant[alive]
        order = np.argsort(tenant, kind="stable")          # tenant-sorted, host order kept inside a tenant
        perm = alive[order]
        counts = np.bincount(tenant, minlength=len(host.tenants))
        seg = np.zeros(len(host.tenants) + 1, dtype=np.int64)
        np.cumsum(counts, out=seg[1:])
        self._res_of[perm] = np.arange(n, dtype=np.int64)
        cap = max(n + 4096, int(n * 1.25))
        self._perm = np.full(cap, -1, dtype=np.int64)       # sized like the capacity buffer: appends are O(new rows)
        self._perm[:n] = perm
        G = len(self.devices)
        # row-range shards of the tenant-sorted order, balanced by rows; the last shard carries the append slack
        self._bounds = [int(round(g * n / G)) for g in range(G)] + [n]
        self._bufs = [ResidentIndex.alloc_rows((self._bounds[g + 1] - self._bounds[g]) if g < G - 1 else cap - self._bounds[g],
                                               host.dim, self.dtype, self.devices[g]) for g in range(G)]
        self._buf = self._bufs[-1]
        t0 = time.perf_counter()
        moved = 0
        jobs = []                                           # (shard, source rows [i0, i1) of a host block, local row map)
        for blk in host.blocks:                             # each host block is read in place, once per shard it lands in
            dst = self._res_of[blk.start:blk.start + blk.n]
            if G == 1:
                if (dst >= 0).any():
                    jobs.append((0, blk.emb, dst))
            else:
                for g in range(G):
                    lo, hi = self._bounds[g], self._bounds[g + 1]
                    hit = np.nonzero((dst >= lo) & (dst < hi))[0]
                    if hit.size == 0:
                        continue
                    i0, i1 = int(hit[0]), int(hit[-1]) + 1     # only the slice of the block that holds this shard's rows
                    sub = dst[i0:i1]
                    jobs.append((g, blk.emb[i0:i1], np.where((sub >= lo) & (sub < hi), sub - lo, -1)))
            moved += blk.n
        if G == 1:
            for g, src, local in jobs:
                ResidentIndex.load_rows_into(self._bufs[g], src, dst_rows=local)
        else:
            # one uploader thread per device: every GPU has its own PCIe link, the C call releases the GIL
            from concurrent.futures import ThreadPoolExecutor

            def upload(g):
                for gg, src, local in jobs:
                    if gg == g:
                        ResidentIndex.load_rows_into(self._bufs[g], src, dst_rows=local)

            with ThreadPoolExecutor(G) as pool:
                list(pool.map(upload, range(G)))
        dt = time.perf_counter() - t0
        self.last_load_gbs = moved * host.dim * 4 / dt / 1e9 if dt > 0 else None
        self._n_res = n
        self._shards = []
        for g in range(G):
            ng = self._bounds[g + 1] - self._bounds[g]
            sh = ResidentIndex(self._bufs[g][:ng], row_base=self._bounds[g])
            sh.update(self._bufs[g], ng)
            sh.set_query_precision(self.query_precision)
            self._shards.append(sh)
        self._resident = self._shards[-1]
        self._multi = MultiIndex(self._shards) if G > 1 else None
        self._ranges = {host.tenants[i]: [[int(seg[i]), int(seg[i + 1])]] for i in range(len(host.tenants)) if counts[i]}
        self.rebuilds += 1

    def _pending_rows(self) -> np.ndarray:
        lo, hi = self._synced_upto, self.host.n_total
        return np.nonzero(self.host.alive[lo:hi])[0] + lo

    def _apply_deltas(self) -> None:
        """Tombstone replaced rows and append the pending rows as per-tenant delta segments."""
        if self._pending_tomb:
            tomb = np.asarray(self._pending_tomb, dtype=np.int64)
            G = len(self._bufs)
            which = np.minimum(np.searchsorted(np.asarray(self._bounds[:G]), tomb, side="right") - 1, G - 1)
            for g in np.unique(which):
                idx = torch.as_tensor(tomb[which == g] - self._bounds[int(g)], dtype=torch.int64, device=self._bufs[int(g)].device)
                self._bufs[int(g)].index_fill_(0, idx, float("nan"))
            self._tomb += len(self._pending_tomb)
            self._pending_tomb = []
        rows = self._pending_rows()
        self._synced_upto = self.host.n_total
        if rows.size:
            tenant = self.host.tenant[rows]
            order = np.argsort(tenant, kind="stable")      # group the batch by tenant, arrival order inside
            rows, tenant = rows[order], tenant[order]
            m = int(rows.size)
            lo = self._n_res
            ResidentIndex.load_rows_into(self._buf, self.host.gather(rows), dst_offset=lo - self._bounds[-2])
            self._perm[lo:lo + m] = rows
            self._res_of[rows] = lo + np.arange(m, dtype=np.int64)
            cuts = np.nonzero(np.diff(tenant))[0] + 1
            for start, end in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [m]])):
                rl = self._ranges.setdefault(self.host.tenants[int(tenant[start])], [])
                if rl and rl[-1][1] == lo + int(start):
                    rl[-1][1] = lo + int(end)                 # contiguous with the tenant's last range
                else:
                    rl.append([lo + int(start), lo + int(end)])
            self._n_res = lo + m
            self.appends += 1
        self._resident.update(self._buf, self._n_res - self._bounds[-2])   # (the last shard when sharded)
        self._seg_key = None

    def resident(self) -> Optional[ResidentIndex]:
        with self._lock:
            dirty = self._synced_upto != self.host.n_total or self._pending_tomb
            if not self._force_rebuild and not dirty:
                return self._resident
            n_new = int(self.host.alive[self._synced_upto:self.host.n_total].sum())
            need_rebuild = (
                self._force_rebuild or self._resident is None or self._buf is None
                or self._n_res + n_new > self._bounds[-2] + self._buf.shape[0]
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
        with self._lock:
            res = self.resident()
            if res is None:
                return None
            ranges = [self._ranges.get(str(u)) or [] for u in user_ids]
            on_device = isinstance(vectors, torch.Tensor)     # e.g. straight out of a device encoder: no host hop
            if self._multi is not None:       # sharded: merged host result -> the collector device (where K5 runs)
                host_q = vectors.detach().float().cpu().numpy() if on_device else vectors
                s, r = self._multi.search_host(np.ascontiguousarray(host_q, dtype=np.float32), limit, ranges)
                return torch.from_numpy(s).to(self.devices[0]), torch.from_numpy(r).to(self.devices[0])
            if on_device:
                q = vectors.to(device=self.device, dtype=torch.float32).contiguous()
            else:
                q = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32)).to(self.device)
            return res.search_ranges(q, limit, ranges)

    def row_identity(self, resident_row: int) -> int:
        return int(self._perm[resident_row])

    def chunk_id_of(self, resident_row: int) -> str:
        return self.host.chunk_id_at(int(self._perm[resident_row]))

    def search(self, user_ids: Sequence[str], vectors: np.ndarray, top_k: int) -> List[List[Dict[str, Any]]]:
        limit = max(int(top_k), 1)
        if limit > N.MMR_MAX_K:
            raise N.NativeError(f"top_k {limit} > {N.MMR_MAX_K}: not supported by the resident-index kernels")
        with self._lock:
            res = self.resident()
            out: List[List[Dict[str, Any]]] = [[] for _ in user_ids]
            if res is None:
                return out
            ranges = [self._ranges.get(str(u)) for u in user_ids]
            live = [i for i, r in enumerate(ranges) if r]
            if not live:
                return out
            q = vectors if len(live) == len(ranges) else vectors[live]
            if q.dtype != np.float32 or not q.flags.c_contiguous:
                q = np.ascontiguousarray(q, dtype=np.float32)
            first = ranges[live[0]]
            if self._multi is not None:
                # row-range shards on the GPUs of this box: one launch per device, fused exchange, mapped mailbox
                scores, rows = self._multi.search_host(q, limit, [ranges[i] for i in live])
            elif len(first) == 1 and all(ranges[i] == first for i in live):
                # one shared range: host-buffer C call (H2D + scan + D2H inside the library)
                lo, hi = first[0]
                if self._seg_key != (lo, hi, self._n_res):          # re-point the one-segment table only when it changes
                    res.update(self._buf, self._n_res, seg_offsets=[lo, hi])
                    self._seg_key = (lo, hi, self._n_res)
                if len(live) == 1 and hasattr(res, "search_host_one"):
                    scores, rows = res.search_host_one(q, limit, 0)      # cached outputs: consumed below, under the lock
                else:
                    scores, rows = res.search_host(q, limit, [0] * len(live))
            else:
                s_dev, r_dev = res.search_ranges(torch.from_numpy(q).to(self.device), limit, [ranges[i] for i in live])
                scores, rows = s_dev.cpu().numpy(), r_dev.cpu().numpy()
            # Lance returns the f32 cosine distance; _format_results (:130-131) turns it into a Python float score
            one = np.float32(1.0)
            sims = (1.0 - (one - scores).astype(np.float64)).tolist()
            hit_rows = rows.tolist()
            host, perm = self.host, self._perm
            for j, i in enumerate(live):
                rr = hit_rows[j]
                n = len(rr) if rr[-1] >= 0 else rr.index(-1)      # hits are a prefix of the result row
                hrows = [perm[r] for r in rr[:n]]
                ids, metas = host.hits_at(hrows)
                sj = sims[j]
                out[i] = [{"chunk_id": ids[p], "score": sj[p], "meta": metas[p]} for p in range(n)]
            return out


class B200Store:
    """Drop-in for `app.ml.retrieve._LANCEDB_STORE` / `app.ml.index_build._LANCEDB_STORE`.

    `db_path` (the reference's LanceDB directory, lancedb_store.py:27-29) makes the store durable and shared: every
    upsert batch is written as an Arrow delta file before `index_versions.json` is bumped, and a store in ANOTHER
    process on the same directory sees it on its next call (it polls the manifest stamps; reference deployment:
    Celery writer + API reader, docker-compose.yml:36-45).  Without `db_path` the store lives in memory only.
    """

    def __init__(self, db_path: Optional[str] = None, device: Any = "cuda:0", dtype: str = "bf16",
                 devices: Optional[Sequence[Any]] = None, query_precision: str = "rescore") -> None:
        """devices=[0, 1, ..., 7]: row-range-shard every collection over those GPUs of this box (one process; the same
        search_* calls; results bit-identical to one GPU).  query_precision: see _Collection."""
        N.lib()  # fail now, loudly, if the CUDA library is missing
        if not torch.cuda.is_available():
            raise N.NativeError("no CUDA device: B200Store has no CPU fallback")
        devs = [torch.device("cuda", d) if isinstance(d, int) else torch.device(d) for d in devices] if devices else None
        self._init_state(db_path, devs[0] if devs else torch.device(device), dtype, devs, query_precision)

    def _init_state(self, db_path: Optional[str], device: torch.device, dtype: str, devices=None,
                    query_precision: str = "rescore") -> None:
        self._device = device
        self._db_path = db_path
        if db_path:
            os.makedirs(db_path, exist_ok=True)

        def log(name):
            return DurableLog(db_path, name) if db_path else None

        self._text_table = _Collection("text_collection", device, dtype, log("text_collection"), query_precision, devices)
        self._image_table = _Collection("image_collection", device, dtype, log("image_collection"), query_precision, devices)
        self._versions = VersionFile(os.path.join(db_path, "index_versions.json") if db_path else None)
        self._sync()   # durable state next to index_versions.json: a restarted process reloads it into HBM

    def _sync(self) -> None:
        """Cross-process invalidation: pick up what other processes made durable (two os.stat calls when idle)."""
        self._text_table.refresh()
        self._image_table.refresh()

    def _collection(self, name: str) -> _Collection:
        return {"text_collection": self._text_table, "image_collection": self._image_table}[name]

    def load_arrow_ipc(self, collection: str, path: str) -> None:
        """Bulk-load an Arrow IPC file written with the reference schema (see make_arrow_table / persist)."""
        import pyarrow as pa
        import pyarrow.ipc as ipc

        self.load_arrow(collection, ipc.open_file(pa.memory_map(path, "r")).read_all())

    def persist(self) -> None:
        """Durable compaction of both collections (alive rows -> new base generation, delta files dropped)."""
        if not self._db_path:
            raise ValueError("B200Store was created without a db_path")
        for coll in (self._text_table, self._image_table):
            coll.persist()

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
        embedding: list<float32>, meta) -- what `lancedb.Table.to_arrow()` yields.  The embedding column is used in
        place (zero-copy) as the host master copy."""
        for user in self._collection(collection).add_table(table):
            self._versions.bump(user)

    # reads (lancedb_store.py:103-123)
    def search_text(self, user_id: str, query_vec: Sequence[float], top_k: int) -> List[Dict[str, Any]]:
        self._text_table.refresh()
        return self._text_table.search([user_id], _query_row(query_vec), top_k)[0]

    def search_image(self, user_id: str, query_vec: Sequence[float], top_k: int) -> List[Dict[str, Any]]:
        self._image_table.refresh()
        return self._image_table.search([user_id], _query_row(query_vec), top_k)[0]

    # micro-batched requests: one launch for B (tenant, query) pairs
    def search_text_batch(self, user_ids: Sequence[str], query_vecs, top_k: int) -> List[List[Dict[str, Any]]]:
        self._text_table.refresh()
        return self._text_table.search(list(user_ids), np.asarray(query_vecs, dtype=np.float32), top_k)

    def search_image_batch(self, user_ids: Sequence[str], query_vecs, top_k: int) -> List[List[Dict[str, Any]]]:
        self._image_table.refresh()
        return self._image_table.search(list(user_ids), np.asarray(query_vecs, dtype=np.float32), top_k)

    def fused_search_batch(self, user_ids: Sequence[str], text_vecs, image_vecs, top_k_text: int, top_k_image: int,
                           final_n: int, tau: float):
        """Rerank-off request path kept on the device: text scan + image scan + z-score fusion + FINAL_N cut +
        CONFIDENCE_TAU gate (kernels K1/K2 + K5).  Returns per request ([{"chunk_id", "modality", "score",
        "combined_score"}, ...] best first, low_confidence flag) -- what retrieve() then _confidence_low() give when
        every hit survives the metadata join (reference app/ml/retrieve.py:103-117, app/ml/generate.py:56-60)."""
        self._sync()
        return _fused_batch(self, list(user_ids), text_vecs, image_vecs, top_k_text, top_k_image, final_n, tau)

    def fused_search_batch_rerank(self, user_ids: Sequence[str], queries: Sequence[str], text_vecs, image_vecs,
                                  top_k_text: int, top_k_image: int, rerank_topk: int, final_n: int, tau: float,
                                  cross_encoder, metadata_store):
        """Rerank-ON request path kept on the device: text scan + image scan, ONE cross-encoder forward
        pass over the (query, passage) pairs of the whole micro-batch (<= RERANK_TOPK per request, reference
        app/ml/retrieve.py:141-148), logits left on the device, then mmr_fuse_f64 = _rerank_text's re-ordering +
        _fuse_results + _confidence_low.  Returns per request (items, low_confidence) like fused_search_batch -- items of
        reranked text hits also carry "rerank_score" -- or None for a request whose candidates could not be joined with
        their text (the caller then serves it through the host path)."""
        self._sync()
        return _fused_batch_rerank(self, list(user_ids), list(queries), text_vecs, image_vecs, top_k_text, top_k_image,
                                   rerank_topk, final_n, tau, cross_encoder, metadata_store)


def _query_row(query_vec) -> np.ndarray:
    """One query as a [1, D] float32 array.  The reference passes a Python list of floats (retrieve.py:53,84): array('f')
    narrows it in C in ~60 % of the time np.asarray(list, float32) takes (same round-to-nearest values)."""
    if isinstance(query_vec, (list, tuple)):
        try:
            return np.frombuffer(array("f", query_vec), dtype=np.float32)[None, :]
        except (TypeError, OverflowError):
            pass
    return np.asarray(query_vec, dtype=np.float32)[None, :]


def _as_queries(v):
    return v if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float32)


def _fused_batch_rerank(store: "B200Store", user_ids, queries, text_vecs, image_vecs, kt: int, ki: int, rerank_topk: int,
                        final_n: int, tau: float, cross_encoder, metadata_store):
    from .index import fuse_f64

    n = len(user_ids)
    t = store._text_table.search_device(user_ids, _as_queries(text_vecs), kt)
    i = store._image_table.search_device(user_ids, _as_queries(image_vecs), ki)
    if t is None and i is None:
        return [([], True) for _ in user_ids]
    dev = (t if t is not None else i)[0].device
    one = torch.tensor(1.0, dtype=torch.float32, device=dev)

    def py_scores(pair, k):      # _format_results: score = 1.0 - float(f32 distance); counts = hits per request
        if pair is None:
            return torch.zeros((n, max(k, 1)), dtype=torch.float64, device=dev), torch.zeros(n, dtype=torch.int32, device=dev), None
        s, r = pair
        return 1.0 - (one - s).double(), (r >= 0).sum(dim=1).to(torch.int32), r

    ts, tc, tr = py_scores(t, kt)
    is_, ic, ir = py_scores(i, ki)
    # the cross-encoder needs the passages of the first RERANK_TOPK text hits of every request: one batched join
    head = min(rerank_topk, ts.shape[1])
    tr_host = tr[:, :head].cpu().numpy() if tr is not None else np.full((n, 0), -1)
    cand_ids = [[store._text_table.chunk_id_of(int(r)) for r in row if r >= 0] for row in tr_host]
    flat = [c for row in cand_ids for c in row]
    bulk = getattr(metadata_store, "get_chunks", None)
    found = bulk(flat) if (bulk and flat) else {c: metadata_store.get_chunk(c) for c in flat}
    pairs, owner, broken = [], [], set()
    for b, row in enumerate(cand_ids):
        for c in row:
            chunk = found.get(c)
            if not chunk or not chunk.text:
                broken.add(b)
                break
        if b not in broken:
            for c in row:
                pairs.append((queries[b], found[c].text))
                owner.append(b)
    rr = torch.zeros((n, ts.shape[1]), dtype=torch.float64, device=dev)
    rc = torch.zeros(n, dtype=torch.int32, device=dev)
    if pairs:
        logits = cross_encoder.predict_device(pairs).to(dev).double()          # ONE forward pass for the micro-batch
        counts = np.bincount(np.asarray(owner), minlength=n)
        rows_idx = torch.as_tensor(owner, device=dev, dtype=torch.int64)
        pos = torch.as_tensor(np.concatenate([np.arange(c) for c in counts]) if len(owner) else [], device=dev, dtype=torch.int64)
        rr[rows_idx, pos] = logits
        rc = torch.as_tensor(counts, dtype=torch.int32, device=dev)
    out = fuse_f64(ts if t is not None else None, tc if t is not None else None, is_ if i is not None else None,
                   ic if i is not None else None, final_n, tau, rerank=rr if t is not None else None,
                   rerank_count=rc if t is not None else None)
    comb, index, low = out["combined"].cpu().numpy(), out["index"].cpu().numpy(), out["low_conf"].cpu().numpy()
    ts_h = ts.cpu().numpy() if t is not None else None
    is_h = is_.cpu().numpy() if i is not None else None
    tr_all = tr.cpu().numpy() if tr is not None else None
    ir_all = ir.cpu().numpy() if ir is not None else None
    rr_h, rc_h = rr.cpu().numpy(), rc.cpu().numpy()
    kt_eff = ts.shape[1] if t is not None else 0
    results = []
    for b in range(n):
        if b in broken:
            results.append(None)
            continue
        items = []
        for o in range(final_n):
            ix = int(index[b, o])
            if ix < 0:
                break
            if ix < kt_eff and t is not None:
                it = {"chunk_id": store._text_table.chunk_id_of(int(tr_all[b, ix])), "modality": "text",
                      "score": float(ts_h[b, ix])}
                if ix < rc_h[b]:
                    it["rerank_score"] = float(rr_h[b, ix])
            else:
                j = ix - kt_eff
                it = {"chunk_id": store._image_table.chunk_id_of(int(ir_all[b, j])), "modality": "image", "score": float(is_h[b, j])}
            it["combined_score"] = float(comb[b, o])
            items.append(it)
        results.append((items, bool(low[b])))
    return results


def _fused_batch(store: "B200Store", user_ids, text_vecs, image_vecs, kt: int, ki: int, final_n: int, tau: float):
    """scan(text) + scan(image) + K5 fusion/gate for a batch of requests, results fetched with one small D2H."""
    from .index import fuse

    t = store._text_table.search_device(user_ids, _as_queries(text_vecs), kt)
    i = store._image_table.search_device(user_ids, _as_queries(image_vecs), ki)
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
            items.append({"chunk_id": coll.chunk_id_of(int(rows[b, o])), "modality": "text" if mod[b, o] == 0 else "image",
                          "score": float(score[b, o]), "combined_score": float(comb[b, o])})
        results.append((items, bool(low[b])))
    return results
