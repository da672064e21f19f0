"""CPU restatement (numpy, float32) of the reference's vector-store search path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Never imported by the
product package.

Follows, function by function:
  * ``normalize``        <- LanceDBStore._normalize        app/storage/lancedb_store.py:63-69
  * ``where_clause``     <- LanceDBStore._where_clause     app/storage/lancedb_store.py:141-144
  * ``prepare_rows``     <- LanceDBStore._prepare_rows     app/storage/lancedb_store.py:71-85
  * ``flat_search``      <- table.search(v).where(user).metric("cosine").limit(max(k,1)).to_list()
                            app/storage/lancedb_store.py:105-111 / 116-122
                            [ext] Lance flat KNN, cosine metric: d = 1 - x.q/(|x||q|) in f32,
                            k smallest d.  PARITY UNPINNED (lancedb not installed, no version pin).
  * ``format_results``   <- LanceDBStore._format_results   app/storage/lancedb_store.py:125-139
  * ``OracleStore``      <- LanceDBStore.search_text/search_image/upsert_*  :87-123

Tie rule (the reference leaves it to Lance; we fix it so every path is
deterministic): total order (distance ascending, row ordinal ascending); the
stable descending sort of ``format_results`` then preserves it.

Filter semantics: prefilter (exact top-k *within* the tenant's rows), SURVEY.md 8(a) a8.

Zero-norm query: Lance would produce 0/0 = NaN distances (unspecified order);
the restatement defines cos = 0 for every row (distance 1.0), which is what the
CUDA path computes.  Documented divergence, unreachable from ``retrieve_*`` unless
the encoder returns an all-zero vector.
"""
from __future__ import annotations

import json
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------- helpers
def normalize(vector: Sequence[float]) -> List[float]:
    """lancedb_store.py:63-69 -- f32 cast, divide by the f32 L2 norm unless it is <= 0."""
    arr = np.asarray(vector, dtype=np.float32)
    norm = np.linalg.norm(arr)
    if norm <= 0:
        return arr.tolist()
    return (arr / norm).tolist()


def normalize_rows(mat: np.ndarray) -> np.ndarray:
    """Row-wise ``normalize`` for an (n, D) matrix, staying in float32 (same arithmetic per row)."""
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    norms = np.sqrt(np.einsum("ij,ij->i", mat, mat, dtype=np.float32)).astype(np.float32)
    safe = np.where(norms > 0, norms, np.float32(1.0)).astype(np.float32)
    return (mat / safe[:, None]).astype(np.float32)


def where_clause(column: str, value: str) -> str:
    """lancedb_store.py:141-144."""
    safe = str(value).replace("'", "''")
    return f"{column} == '{safe}'"


def prepare_rows(rows: Iterable[Any]) -> List[Dict[str, Any]]:
    """lancedb_store.py:71-85 -- rows are objects with the VectorRow attributes."""
    prepared = []
    for row in rows:
        prepared.append(
            {
                "chunk_id": row.chunk_id,
                "user_id": row.user_id,
                "document_id": row.document_id,
                "modality": row.modality,
                "embedding": normalize(row.embedding),
                "meta": json.dumps(row.meta or {}),
            }
        )
    return prepared


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round float32 -> bfloat16 (round-to-nearest-even) and return the value as float32.

    Used by the *strict* parity mode: the oracle evaluated on exactly the values the
    resident bf16 index holds must give identical ids.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)


# --------------------------------------------------------------------------- the scan
def cosine_distances(rows: np.ndarray, q: np.ndarray, unit_rows: bool = False) -> np.ndarray:
    """f32 cosine distance of ``q`` to every row: d = 1 - x.q/(|x||q|)  ([ext] Lance `cosine`).

    ``unit_rows=True`` skips the division by |x| (rows were written through
    ``normalize``, lancedb_store.py:81, so |x| = 1 to f32 rounding; the difference is <= ~1e-7).
    """
    rows = np.asarray(rows, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    dots = rows @ q  # f32 accumulate (OpenBLAS sgemv)
    qn = np.float32(np.sqrt(np.dot(q, q)))
    if qn <= 0:
        return np.ones(rows.shape[0], dtype=np.float32)
    if unit_rows:
        cos = dots / qn
    else:
        xn = np.sqrt(np.einsum("ij,ij->i", rows, rows, dtype=np.float32)).astype(np.float32)
        denom = xn * qn
        cos = np.where(denom > 0, dots / np.where(denom > 0, denom, np.float32(1)), np.float32(0))
    return (np.float32(1.0) - cos.astype(np.float32)).astype(np.float32)


def topk_smallest(dist: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k smallest distances under (distance asc, ordinal asc)."""
    n = dist.shape[0]
    k = min(k, n)
    if k <= 0:
        return np.empty(0, dtype=np.int64)
    if n > 4 * k + 64:
        kth = np.partition(dist, k - 1)[k - 1]
        cand = np.nonzero(dist <= kth)[0]
    else:
        cand = np.arange(n)
    order = np.argsort(dist[cand], kind="stable")  # ties keep ascending ordinal
    return cand[order[:k]].astype(np.int64)


def flat_search(
    rows: np.ndarray,
    q: Sequence[float],
    k: int,
    lo: int = 0,
    hi: Optional[int] = None,
    unit_rows: bool = True,
) -> Tuple[np.ndarray, np.ndarray]:
    """Exact flat KNN of one query over rows[lo:hi]; returns (distance f32[k'], global ordinal i64[k']).

    ``q`` is re-normalised first exactly as search_text does (lancedb_store.py:104);
    ``k`` is clamped with max(k, 1) (:109).
    """
    hi = rows.shape[0] if hi is None else hi
    k = max(int(k), 1)
    if hi <= lo:
        return np.empty(0, np.float32), np.empty(0, np.int64)
    qn = np.asarray(normalize(q), dtype=np.float32)
    dist = cosine_distances(rows[lo:hi], qn, unit_rows=unit_rows)
    ids = topk_smallest(dist, k)
    return dist[ids], ids + lo


def flat_search_batch(
    rows: np.ndarray,
    queries: np.ndarray,
    k: int,
    lo: int = 0,
    hi: Optional[int] = None,
    block: int = 262144,
) -> Tuple[np.ndarray, np.ndarray]:
    """Batched ``flat_search`` (unit rows): one sgemm per row block, running top-k per query.

    Returns (distance f32[B,k'], ordinal i64[B,k']) with k' = min(max(k,1), hi-lo).
    Identical ordering rule as ``flat_search``.
    """
    hi = rows.shape[0] if hi is None else hi
    k = max(int(k), 1)
    queries = np.atleast_2d(np.asarray(queries, dtype=np.float32))
    b = queries.shape[0]
    qn = np.stack([np.asarray(normalize(qv), dtype=np.float32) for qv in queries])
    kk = min(k, max(hi - lo, 0))
    best_d = np.empty((b, 0), np.float32)
    best_i = np.empty((b, 0), np.int64)
    for s in range(lo, hi, block):
        e = min(s + block, hi)
        d = (np.float32(1.0) - (rows[s:e] @ qn.T).astype(np.float32)).T  # [b, e-s]
        zero_q = ~(np.einsum("ij,ij->i", qn, qn) > 0)
        if zero_q.any():
            d[zero_q] = np.float32(1.0)
        kb = min(kk, e - s)
        if e - s > 4 * kb + 64:
            part = np.argpartition(d, kb - 1, axis=1)[:, :kb]
            kth = np.take_along_axis(d, part, axis=1).max(axis=1)
        else:
            kth = np.full(b, np.inf, np.float32)
        cd_list, ci_list = [], []
        for qi in range(b):
            cand = np.nonzero(d[qi] <= kth[qi])[0]
            o = np.argsort(d[qi][cand], kind="stable")[:kb]
            cd_list.append(d[qi][cand[o]])
            ci_list.append(cand[o].astype(np.int64) + s)
        cd = np.stack(cd_list)
        ci = np.stack(ci_list)
        alld = np.concatenate([best_d, cd], axis=1)
        alli = np.concatenate([best_i, ci], axis=1)
        # merge: earlier blocks have smaller ordinals, concatenation order is ordinal order per tie
        o = np.argsort(alld, axis=1, kind="stable")[:, :kk]
        best_d = np.take_along_axis(alld, o, axis=1)
        best_i = np.take_along_axis(alli, o, axis=1)
    return best_d, best_i


def flat_search_threads(
    rows: np.ndarray,
    queries: np.ndarray,
    k: int,
    threads: int,
    lo: int = 0,
    hi: Optional[int] = None,
) -> Tuple[np.ndarray, np.ndarray]:
    """``flat_search_batch`` with the row range cut into ``threads`` contiguous slices scanned concurrently
    (numpy releases the GIL inside the sgemm and the partition), then one stable merge in slice order.

    Same results as ``flat_search_batch`` (same tie rule: slices are concatenated in ordinal order).  This is the
    "all host cores" form of the CPU baseline in bench.py: Lance's flat scan is multi-threaded too.
    """
    from concurrent.futures import ThreadPoolExecutor

    hi = rows.shape[0] if hi is None else hi
    k = max(int(k), 1)
    queries = np.atleast_2d(np.asarray(queries, dtype=np.float32))
    n = max(hi - lo, 0)
    threads = max(1, min(int(threads), (n + 65535) // 65536 or 1))
    if threads == 1:
        return flat_search_batch(rows, queries, k, lo, hi)
    cuts = [lo + (n * t) // threads for t in range(threads + 1)]
    with ThreadPoolExecutor(threads) as pool:
        parts = list(pool.map(lambda t: flat_search_batch(rows, queries, k, cuts[t], cuts[t + 1]), range(threads)))
    alld = np.concatenate([p[0] for p in parts], axis=1)
    alli = np.concatenate([p[1] for p in parts], axis=1)
    kk = min(k, n)
    o = np.argsort(alld, axis=1, kind="stable")[:, :kk]
    return np.take_along_axis(alld, o, axis=1), np.take_along_axis(alli, o, axis=1)


# --------------------------------------------------------------------------- result shaping
def format_results(rows: List[Dict[str, Any]]) -> List[Dict[str, Any]]:
    """lancedb_store.py:125-139 -- similarity = 1.0 - float(_distance); stable sort by score desc."""
    formatted = []
    for row in rows:
        distance = float(row.get("_distance", 0.0))
        similarity = 1.0 - distance
        formatted.append(
            {
                "chunk_id": row.get("chunk_id"),
                "score": similarity,
                "meta": json.loads(row.get("meta") or "{}"),
            }
        )
    formatted.sort(key=lambda item: item["score"], reverse=True)
    return formatted


class OracleTable:
    """One shared collection (text_collection / image_collection) held as numpy columns."""

    def __init__(self) -> None:
        self.chunk_id: List[str] = []
        self.user_id: List[str] = []
        self.meta: List[Optional[str]] = []
        self.embedding: List[np.ndarray] = []

    def upsert(self, payloads: List[Dict[str, Any]]) -> None:
        """delete-by-chunk_id then add (lancedb_store.py:91-93)."""
        if not payloads:
            return
        drop = {p["chunk_id"] for p in payloads}
        keep = [i for i, c in enumerate(self.chunk_id) if c not in drop]
        self.chunk_id = [self.chunk_id[i] for i in keep]
        self.user_id = [self.user_id[i] for i in keep]
        self.meta = [self.meta[i] for i in keep]
        self.embedding = [self.embedding[i] for i in keep]
        for p in payloads:
            self.chunk_id.append(p["chunk_id"])
            self.user_id.append(p["user_id"])
            self.meta.append(p["meta"])
            self.embedding.append(np.asarray(p["embedding"], dtype=np.float32))

    def search(self, user_id: str, vector: Sequence[float], limit: int) -> List[Dict[str, Any]]:
        """search(v).where(user_id == ..).metric('cosine').limit(n).to_list() with prefilter."""
        sel = [i for i, u in enumerate(self.user_id) if u == str(user_id)]
        if not sel:
            return []
        mat = np.stack([self.embedding[i] for i in sel])
        dist = cosine_distances(mat, np.asarray(vector, dtype=np.float32), unit_rows=False)
        ids = topk_smallest(dist, limit)
        return [
            {"chunk_id": self.chunk_id[sel[j]], "meta": self.meta[sel[j]], "_distance": dist[j]}
            for j in ids
        ]


class OracleStore:
    """Duck-type of LanceDBStore (search_text/search_image/upsert_*), lancedb_store.py:24-123."""

    def __init__(self) -> None:
        self._text_table = OracleTable()
        self._image_table = OracleTable()

    def upsert_text_vectors(self, rows: Iterable[Any]) -> None:
        self._text_table.upsert(prepare_rows(rows))

    def upsert_image_vectors(self, rows: Iterable[Any]) -> None:
        self._image_table.upsert(prepare_rows(rows))

    def search_text(self, user_id: str, query_vec: Sequence[float], top_k: int) -> List[Dict[str, Any]]:
        vector = normalize(query_vec)
        return format_results(self._text_table.search(user_id, vector, max(top_k, 1)))

    def search_image(self, user_id: str, query_vec: Sequence[float], top_k: int) -> List[Dict[str, Any]]:
        vector = normalize(query_vec)
        return format_results(self._image_table.search(user_id, vector, max(top_k, 1)))
