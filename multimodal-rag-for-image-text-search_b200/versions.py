"""Per-user index version: the invalidation key of the result caches and of the resident index.
Mirrors get_index_version / _bump_version (reference app/ml/index_build.py:33-43): a JSON dict
{user_id: int} in `index_versions.json`, bumped on every upsert.  Reads are cached on the file's
(mtime_ns, size) so a request does not re-parse the file three times (SURVEY 3.2 hot loop 5)."""
from __future__ import annotations

import json
import os
import threading
from typing import Dict, Optional, Tuple


class VersionFile:
    def __init__(self, path: Optional[str]) -> None:
        self._path = path
        self._lock = threading.Lock()
        self._mem: Dict[str, int] = {}
        self._stamp: Optional[Tuple[int, int]] = None

    def _load(self) -> Dict[str, int]:
        if self._path is None:
            return self._mem
        try:
            st = os.stat(self._path)
        except OSError:
            self._mem, self._stamp = {}, None
            return self._mem
        stamp = (st.st_mtime_ns, st.st_size)
        if stamp != self._stamp:
            try:
                with open(self._path) as fh:
                    data = json.load(fh)
                self._mem = {str(k): int(v) for k, v in data.items()}
            except Exception:
                self._mem = {}
            self._stamp = stamp
        return self._mem

    def get(self, user_id: str) -> int:
        with self._lock:
            return self._load().get(str(user_id), 0)

    def bump(self, user_id: str) -> int:
        with self._lock:
            cur = dict(self._load())
            cur[str(user_id)] = cur.get(str(user_id), 0) + 1
            self._mem = cur
            if self._path is not None:
                os.makedirs(os.path.dirname(os.path.abspath(self._path)), exist_ok=True)
                tmp = self._path + ".tmp"
                with open(tmp, "w") as fh:
                    json.dump(cur, fh)
                os.replace(tmp, self._path)
                st = os.stat(self._path)
                self._stamp = (st.st_mtime_ns, st.st_size)
            return cur[str(user_id)]
