"""TTL caches that sit above the scan, same keys and lifetimes as the reference (app/cache/__init__.py):
query embeddings 300 s keyed by the normalised query (:44-58); retrieval results 120 s keyed by
(user_id, normalised query, index version) (:64-80)."""
from __future__ import annotations

import time
from typing import Any, Dict, Hashable, Optional, Tuple

EMBED_TTL_SEC = 300
RETRIEVAL_TTL_SEC = 120


def normalize_query(query: str) -> str:
    return " ".join(query.strip().lower().split())


class TTLCache:
    def __init__(self) -> None:
        self._data: Dict[Hashable, Tuple[float, Any]] = {}

    def get(self, key: Hashable) -> Optional[Any]:
        hit = self._data.get(key)
        if hit is None:
            return None
        if hit[0] < time.time():
            self._data.pop(key, None)
            return None
        return hit[1]

    def put(self, key: Hashable, value: Any, ttl: float) -> None:
        self._data[key] = (time.time() + ttl, value)

    def clear(self) -> None:
        self._data.clear()


_EMBED = TTLCache()
_RETRIEVAL = TTLCache()


def clear_all_caches() -> None:
    _EMBED.clear()
    _RETRIEVAL.clear()


def get_query_embeddings(query: str):
    return _EMBED.get(normalize_query(query))


def set_query_embeddings(query: str, text_vec, image_vec, ttl: int = EMBED_TTL_SEC) -> None:
    _EMBED.put(normalize_query(query), (text_vec, image_vec), ttl)


def get_retrieval_results(user_id: str, query: str, index_version: int):
    return _RETRIEVAL.get((user_id, normalize_query(query), index_version))


def set_retrieval_results(user_id: str, query: str, index_version: int, results, ttl: int = RETRIEVAL_TTL_SEC) -> None:
    _RETRIEVAL.put((user_id, normalize_query(query), index_version), results, ttl)
