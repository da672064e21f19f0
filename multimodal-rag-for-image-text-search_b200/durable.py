"""Durable state of one collection under `db_path`, shared between processes.

The reference's writer (Celery worker, app/tasks.py:108,165) and reader (API process, app/ml/retrieve.py:21) are
different processes that meet only on disk: the LanceDB directory and `index_versions.json`
(app/ml/index_build.py:33-43).  The resident-index store keeps the same contract with plain Arrow IPC files:

    <name>.manifest.json            {"generation": g, "base": file | null, "deltas": [file, ...]}
    <name>.g<g>.base.arrow          all alive rows at the last compaction (`persist()`), reference schema
    <name>.g<g>.d<seq>.arrow        one file per upsert batch since then (what `table.add` makes durable)
    <name>.lock                     flock() taken by a writer around {refresh, append, manifest update}

A reader polls the manifest's (mtime_ns, size): unchanged -> nothing to do (one os.stat per search); same generation ->
load only the delta files it has not applied; new generation -> reload the base.  The manifest is replaced atomically
(tmp + os.replace), so a reader never sees a half-written state; delta files are complete before the manifest names them.
A directory written by round 1 (`<name>.arrow`, no manifest) is read as generation 0 with no deltas.
"""
from __future__ import annotations

import contextlib
import json
import os
from typing import Dict, List, Optional, Tuple

try:
    import fcntl
except ImportError:  # pragma: no cover - non-POSIX
    fcntl = None


class DurableLog:
    def __init__(self, db_path: str, name: str) -> None:
        self.dir = db_path
        self.name = name
        self.manifest_path = os.path.join(db_path, name + ".manifest.json")
        self.legacy_path = os.path.join(db_path, name + ".arrow")
        self.lock_path = os.path.join(db_path, name + ".lock")

    # ------------------------------------------------------------------------------------------ reading
    def stamp(self) -> Optional[Tuple[int, int, int]]:
        for path in (self.manifest_path, self.legacy_path):
            try:
                st = os.stat(path)
                return (st.st_mtime_ns, st.st_size, st.st_ino)
            except OSError:
                continue
        return None

    def manifest(self) -> Dict:
        try:
            with open(self.manifest_path) as fh:
                m = json.load(fh)
            return {"generation": int(m.get("generation", 0)), "base": m.get("base"), "deltas": list(m.get("deltas", []))}
        except (OSError, ValueError):
            pass
        if os.path.exists(self.legacy_path):
            return {"generation": 0, "base": os.path.basename(self.legacy_path), "deltas": []}
        return {"generation": 0, "base": None, "deltas": []}

    def read_table(self, fname: str):
        import pyarrow as pa
        import pyarrow.ipc as ipc

        # memory-mapped: the embedding column is consumed as a zero-copy view by the loader
        src = pa.memory_map(os.path.join(self.dir, fname), "r")
        return ipc.open_file(src).read_all()

    # ------------------------------------------------------------------------------------------ writing
    @contextlib.contextmanager
    def locked(self):
        os.makedirs(self.dir, exist_ok=True)
        if fcntl is None:
            yield
            return
        fd = os.open(self.lock_path, os.O_CREAT | os.O_RDWR, 0o644)
        try:
            fcntl.flock(fd, fcntl.LOCK_EX)
            yield
        finally:
            fcntl.flock(fd, fcntl.LOCK_UN)
            os.close(fd)

    def _write_table(self, fname: str, table) -> None:
        import pyarrow as pa
        import pyarrow.ipc as ipc

        tmp = os.path.join(self.dir, fname + ".tmp")
        with pa.OSFile(tmp, "wb") as sink, ipc.new_file(sink, table.schema) as writer:
            writer.write_table(table)
        os.replace(tmp, os.path.join(self.dir, fname))

    def _write_manifest(self, m: Dict) -> None:
        tmp = self.manifest_path + ".tmp"
        with open(tmp, "w") as fh:
            json.dump(m, fh)
        os.replace(tmp, self.manifest_path)

    def append_delta(self, table) -> Dict:
        """Make one upsert batch durable; call with the lock held.  Returns the new manifest."""
        m = self.manifest()
        fname = f"{self.name}.g{m['generation']}.d{len(m['deltas']) + 1:06d}.arrow"
        self._write_table(fname, table)
        m["deltas"].append(fname)
        self._write_manifest(m)
        return m

    def write_base(self, table) -> Dict:
        """Compaction: a new generation whose base holds `table` (None = empty), no deltas; old files are removed."""
        old = self.manifest()
        gen = old["generation"] + 1
        base = None
        if table is not None and table.num_rows:
            base = f"{self.name}.g{gen}.base.arrow"
            self._write_table(base, table)
        m = {"generation": gen, "base": base, "deltas": []}
        self._write_manifest(m)
        for fname in [old["base"]] + old["deltas"]:
            if fname and fname != base:
                with contextlib.suppress(OSError):
                    os.remove(os.path.join(self.dir, fname))
        return m
