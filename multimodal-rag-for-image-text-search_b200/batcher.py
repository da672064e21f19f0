"""Micro-batching of concurrent retrieval requests.

Every call in the reference is batch 1 (`chat` -> `retrieve(user_id, query)`, api/routes.py:276): query
batches only arise from concurrent requests.  MicroBatcher is the piece that turns N in-flight requests into ONE
`retrieve_batch_device` call (one text-scan launch + one image-scan launch + K5 for the whole batch): request threads
`submit()` and block on a future; a single worker thread drains the queue into batches of at most `max_batch`, waiting
at most `max_wait_ms` for stragglers after the first request arrives.  Exceptions raised by the batch call propagate
to every request of that batch (the reference lets search errors surface as HTTP 500, app/main.py:39-41).
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import Any, Callable, List, Optional, Sequence, Tuple


class MicroBatcher:
    def __init__(self, serve_batch: Callable[[Sequence[str], Sequence[str]], Sequence[Any]], max_batch: int = 128,
                 max_wait_ms: float = 1.0) -> None:
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self._serve = serve_batch
        self._max_batch = int(max_batch)
        self._max_wait = float(max_wait_ms) / 1e3
        self._q: "queue.Queue[Optional[Tuple[str, str, Future]]]" = queue.Queue()
        self._closed = False
        self.batches: List[int] = []          # size of every batch served (observability / tests)
        self._worker = threading.Thread(target=self._run, name="mmr-microbatcher", daemon=True)
        self._worker.start()

    # -- request side ---------------------------------------------------------------------------
    def submit(self, user_id: str, query: str) -> Future:
        if self._closed:
            raise RuntimeError("MicroBatcher is closed")
        fut: Future = Future()
        self._q.put((user_id, query, fut))
        return fut

    def retrieve(self, user_id: str, query: str, timeout: Optional[float] = None):
        """Blocking per-request call with the shape of the reference's `retrieve(user_id, query)`."""
        return self.submit(user_id, query).result(timeout)

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(None)
            self._worker.join(timeout=5.0)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- worker side ----------------------------------------------------------------------------
    def _collect(self, first) -> Tuple[List[Tuple[str, str, Future]], bool]:
        batch, stop = [first], False
        deadline = time.monotonic() + self._max_wait
        while len(batch) < self._max_batch:
            remaining = deadline - time.monotonic()
            try:
                item = self._q.get_nowait() if remaining <= 0 else self._q.get(timeout=remaining)
            except queue.Empty:
                break
            if item is None:
                stop = True
                break
            batch.append(item)
        return batch, stop

    def _run(self) -> None:
        while True:
            first = self._q.get()
            if first is None:
                return
            batch, stop = self._collect(first)
            live = [(u, q, f) for u, q, f in batch if f.set_running_or_notify_cancel()]
            if live:
                self.batches.append(len(live))
                try:
                    results = self._serve([u for u, _, _ in live], [q for _, q, _ in live])
                    if len(results) != len(live):
                        raise RuntimeError(f"batch call returned {len(results)} results for {len(live)} requests")
                    for (_, _, f), r in zip(live, results):
                        f.set_result(r)
                except BaseException as exc:  # noqa: BLE001 - every waiter must be released
                    for _, _, f in live:
                        if not f.done():
                            f.set_exception(exc)
            if stop:
                return
