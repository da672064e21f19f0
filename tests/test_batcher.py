"""MicroBatcher (host logic, CPU): concurrent requests are served in batches, in order, errors reach every waiter."""
import importlib
import threading
import time

import pytest

PKG = "multimodal-rag-for-image-text-search_b200"
batcher = importlib.import_module(PKG + ".batcher")


def test_concurrent_requests_are_batched_and_answered_individually():
    seen = []

    def serve(users, queries):
        seen.append(list(zip(users, queries)))
        time.sleep(0.01)                        # a "scan" long enough for the next wave to queue up
        return [f"{u}:{q}" for u, q in zip(users, queries)]

    with batcher.MicroBatcher(serve, max_batch=16, max_wait_ms=20.0) as mb:
        out = {}

        def client(i):
            out[i] = mb.retrieve(f"u{i % 3}", f"q{i}")

        threads = [threading.Thread(target=client, args=(i,)) for i in range(40)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(10)
        assert out == {i: f"u{i % 3}:q{i}" for i in range(40)}
        assert sum(mb.batches) == 40 and max(mb.batches) <= 16
        assert len(mb.batches) < 40, "requests in flight together must share launches"
        assert sorted(x for b in seen for x in b) == sorted((f"u{i % 3}", f"q{i}") for i in range(40))


def test_single_request_is_not_held_longer_than_max_wait():
    with batcher.MicroBatcher(lambda u, q: ["ok"] * len(u), max_batch=64, max_wait_ms=5.0) as mb:
        t0 = time.perf_counter()
        assert mb.retrieve("u", "q") == "ok"
        assert time.perf_counter() - t0 < 0.5
        assert mb.batches == [1]


def test_errors_propagate_to_every_request_of_the_batch():
    def serve(users, queries):
        raise ValueError("scan failed")

    with batcher.MicroBatcher(serve, max_batch=8, max_wait_ms=30.0) as mb:
        futs = [mb.submit("u", f"q{i}") for i in range(5)]
        for f in futs:
            with pytest.raises(ValueError, match="scan failed"):
                f.result(5)
        # the worker survives a failed batch
        mb._serve = lambda u, q: ["fine"] * len(u)
        assert mb.retrieve("u", "again", timeout=5) == "fine"


def test_wrong_result_count_and_close():
    mb = batcher.MicroBatcher(lambda u, q: [], max_batch=4, max_wait_ms=1.0)
    with pytest.raises(RuntimeError, match="returned 0 results"):
        mb.retrieve("u", "q", timeout=5)
    mb.close()
    with pytest.raises(RuntimeError, match="closed"):
        mb.submit("u", "q")
    with pytest.raises(ValueError):
        batcher.MicroBatcher(lambda u, q: [], max_batch=0)
