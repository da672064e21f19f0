#!/usr/bin/env python
"""One process, G GPUs behind the drop-in: B200Store(devices=[0..G-1]).search_image on the 10M x 512 table.

    python benchmarks/multi_store_bench.py --gpus 8 [--rows 10000000] [--steps 300]

Prints one JSON line: e2e queries/s through `search_image(user, list, 10)` with the collection row-range-sharded over G
GPUs of this box (single process: per-device launcher threads, fused peer-memory exchange, mapped mailbox), next to the
same store on one GPU, and a bit-identity check between the two."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--k", type=int, default=10)
    args = ap.parse_args()
    import pyarrow as pa
    import torch

    pkg = importlib.import_module("multimodal-rag-for-image-text-search_b200")
    n, d, k = args.rows, 512, args.k
    emb = bench.host_table(n, d, bench.host_cores())
    one = lambda v: pa.array([v], pa.string()).take(pa.array(np.zeros(n, np.int32)))
    table = pkg.make_arrow_table(pa.array(np.arange(n)).cast(pa.string()), one("bench"), one("doc"), one("image"), emb, one("{}"))
    qs = bench.gen_queries(args.steps + 20, d)
    qlists = [q.tolist() for q in qs]
    out = {"workload": f"{n}x{d} bf16, top-{k}, batch 1, B200Store.search_image(user, list, k)", "steps": args.steps}
    results = {}
    for name, devices in (("one_gpu", None), (f"{args.gpus}_gpus_one_process", list(range(args.gpus)))):
        store = pkg.B200Store(devices=devices)
        t0 = time.perf_counter()
        store.load_arrow("image_collection", table)
        store.search_image("bench", qlists[0], k)
        load_s = time.perf_counter() - t0
        for i in range(20):
            store.search_image("bench", qlists[i], k)
        t0 = time.perf_counter()
        for i in range(20, 20 + args.steps):
            hits = store.search_image("bench", qlists[i], k)
        dt = (time.perf_counter() - t0) / args.steps
        # bare C call under it
        res = store._image_table._multi if devices else store._image_table.resident()
        t0 = time.perf_counter()
        for i in range(20, 20 + args.steps):
            res.search_host(qs[i], k)
        dt_c = (time.perf_counter() - t0) / args.steps
        results[name] = [store.search_image("bench", qlists[i], k) for i in range(20, 40)]
        out[name] = {"queries_per_s": 1.0 / dt, "ms_per_query": dt * 1e3, "c_abi_ms_per_query": dt_c * 1e3,
                     "load_and_first_search_s": load_s, "loader_GBs": store._image_table.last_load_gbs}
        del store, res
        torch.cuda.empty_cache()
    a, b = results["one_gpu"], results[f"{args.gpus}_gpus_one_process"]
    out["bit_identical_to_one_gpu"] = a == b
    out["speedup"] = out[f"{args.gpus}_gpus_one_process"]["queries_per_s"] / out["one_gpu"]["queries_per_s"]
    print(json.dumps(out), flush=True)
    assert a == b, "sharded store and single-GPU store disagree"


if __name__ == "__main__":
    main()
