#!/usr/bin/env python
"""Per-call cost of B200Store.search_text on a tiny table (10k x 384: the scan is ~12 us of kernel, the rest is host work):
min / median over repeats of 3000 calls, list-of-floats in, list-of-dicts out; beside it the bare C call with host buffers."""
import importlib
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pkg = importlib.import_module("multimodal-rag-for-image-text-search_b200")


def main():
    n, d = 10_000, 384
    rng = np.random.default_rng(1)
    emb = rng.standard_normal((n, d), dtype=np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    store = pkg.B200Store()
    store.load_arrow("text_collection", pkg.make_arrow_table([f"c{i}" for i in range(n)], ["u"] * n, ["d"] * n, ["text"] * n, emb,
                                                             ["{}"] * n))
    qs = [q.tolist() for q in bench.gen_queries(64, d)]
    qa = bench.gen_queries(64, d)
    res = store._text_table.resident()
    for q in qs:
        store.search_text("u", q, 10)
    e2e, cabi = [], []
    for rep in range(7):
        t0 = time.perf_counter()
        for i in range(3000):
            store.search_text("u", qs[i & 63], 10)
        e2e.append((time.perf_counter() - t0) / 3000 * 1e6)
        t0 = time.perf_counter()
        for i in range(3000):
            res.search_host(qa[i & 63], 10, None)
        cabi.append((time.perf_counter() - t0) / 3000 * 1e6)
    print(json.dumps({"store_search_text_us": {"min": min(e2e), "median": statistics.median(e2e)},
                      "c_call_host_buffers_us": {"min": min(cabi), "median": statistics.median(cabi)},
                      "python_around_the_c_call_us": min(e2e) - min(cabi)}))


if __name__ == "__main__":
    main()
