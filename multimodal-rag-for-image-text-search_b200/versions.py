"""Per-user index version: the invalidation key of the result caches and of the resident index.
Mirrors get_index_version / _bump_version (reference app/ml/index_build.py:33-43): a JSON dict
{user_id: int} in `index_versions.json`, bumped on every upsert.  Reads are cached on the file's
(mtime_ns, size) so a request does not re-parse the file three times (SURVEY 3.2 hot loop 5).  The file is shared
with other processes (the reference's Celery writer and API reader): bumps are read-modify-write under flock(), and a
reader notices another process's bump through the stamp."""
from __future__ import annotations

import contextlib
import json
import os
import threading

try:
    import fcntl
except ImportError:  # pragma: no cover
    fcntl = None
from typing import Dict, Optional, Tuple


class VersionFile:
    def __init__(self, path: Optional[str]) -> None:
        self._path = path
        self._lock = threading.Lock()
        self._mem: Dict[str, int] = {}
        self._stamp: Optional[Tuple[int, int, int]] = None

    def _load(self) -> Dict[str, int]:
        if self._path is None:
            return self._mem
        try:
            st = os.stat(self._path)
        except OSError:
            self._mem, self._stamp = {}, None
            return self._mem
        stamp = (st.st_mtime_ns, st.st_size, st.st_ino)
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

    @contextlib.contextmanager
    def _file_lock(self):
        if self._path is None or fcntl is None:
            yield
            return
        os.makedirs(os.path.dirname(os.path.abspath(self._path)), exist_ok=True)
        fd = os.open(self._path + ".lock", os.O_CREAT | os.O_RDWR, 0o644)
        try:
            fcntl.flock(fd, fcntl.LOCK_EX)
            yield
        finally:
            fcntl.flock(fd, fcntl.LOCK_UN)
            os.close(fd)

    def bump(self, user_id: str) -> int:
        with self._lock, self._file_lock():
            self._stamp = None if self._path is not None else self._stamp   # re-read: another process may have bumped
            cur = dict(self._load())
            cur[str(user_id)] = cur.get(str(user_id), 0) + 1
            self._mem = cur
            if self._path is not None:
                tmp = self._path + ".tmp"
                with open(tmp, "w") as fh:
                    json.dump(cur, fh)
                os.replace(tmp, self._path)
                st = os.stat(self._path)
                self._stamp = (st.st_mtime_ns, st.st_size, st.st_ino)
            return cur[str(user_id)]
