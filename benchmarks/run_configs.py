#!/usr/bin/env python
"""Measure the BASELINE.json configs that bench.py's single headline line does not cover (one GPU).

    python benchmarks/run_configs.py [--configs 1,2,4,5] [--out gpurun_out/configs.json]

C1  all-MiniLM-shaped 10k x 384 text table through the full drop-in (B200Store.search_text, dicts out), top-10
C2  CLIP-shaped 1M x 512 bf16 index, query batch 1..1024, top-10
C4  50M x 384 text + 10M x 512 image, top-50 / top-12, device fusion + CONFIDENCE_TAU gate (rerank off)
C5  1000 tenants of ragged size (1k..1M rows, log-uniform, seed 7) searched in ONE varlen launch
(C3, the 10M x 512 table sharded over 2/4/8 GPUs, is bench.py --gpus N.)
Timing: CUDA events on the launching stream, >= 3 warm-ups; every table is larger than L2 except C1 (stated).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic-table helpers)

PKG = "multimodal-rag-for-image-text-search_b200"
HBM_PEAK, TF_PEAK, PEAK_KIND = bench.measured_peaks()


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def qdev(b, d, seed=1):
    return torch.from_numpy(bench.gen_queries(b, d)).cuda()


def c1(pkg):
    from oracle import flat_search as ofs
    n, d = 10_000, 384
    rng = np.random.default_rng(1)
    emb = rng.standard_normal((n, d), dtype=np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    store = pkg.B200Store()
    store.load_arrow("text_collection", pkg.make_arrow_table([f"c{i}" for i in range(n)], ["u"] * n, ["d"] * n,
                                                             ["text"] * n, emb, ["{}"] * n))
    qs = bench.gen_queries(200, d)
    store.search_text("u", qs[0].tolist(), 10)
    t0 = time.perf_counter()
    for q in qs:
        hits = store.search_text("u", q.tolist(), 10)
    gpu_ms = (time.perf_counter() - t0) / len(qs) * 1e3
    t0 = time.perf_counter()
    for q in qs:
        ofs.flat_search(emb, q, 10)
    cpu_ms = (time.perf_counter() - t0) / len(qs) * 1e3
    # the scan alone through the C ABI with host buffers (H2D + kernel + D2H + sync), no dict building
    res = store._text_table.resident()
    res.search_host(qs[0], 10, None)
    t0 = time.perf_counter()
    for q in qs:
        res.search_host(q, 10, None)
    scan_ms = (time.perf_counter() - t0) / len(qs) * 1e3
    return {"config": "C1 10k x 384 text, top-10, B200Store.search_text end to end (list in, dicts out)",
            "ms_per_query_b200_e2e": gpu_ms, "ms_per_query_b200_scan_only_host_buffers": scan_ms,
            "ms_per_query_cpu_oracle_scan_only": cpu_ms, "hits": len(hits),
            "note": "7.7 MB table: L2-resident, latency-bound (host call + H2D + launch + D2H), not a bandwidth case"}


def sweep(ix, d, k, batches, rows, esize=2):
    out = []
    for b in batches:
        q = qdev(b, d)
        reps = 50 if b <= 256 else 15
        ms = timed(lambda: ix.search(q, k), reps)
        gbs = rows * d * esize / (ms * 1e-3) / 1e9
        tfl = 2.0 * b * rows * d / (ms * 1e-3) / 1e12
        out.append({"batch": b, "ms": ms, "qps": b / (ms * 1e-3), "hbm_GBs": gbs, "hbm_frac": gbs / HBM_PEAK,
                    "tflops": tfl, "tensor_frac": tfl / TF_PEAK})
    return out


def c2(pkg):
    ix = bench.build_shard(pkg, 0, 1_000_000, 512, "bf16", torch.device("cuda:0"))
    res = sweep(ix, 512, 10, [1, 2, 4, 8, 16, 64, 128, 256, 1024], 1_000_000)
    ix.close()
    return {"config": "C2 1M x 512 bf16, top-10, query batch 1..1024 (1.02 GB table > 126 MB L2)", "sweep": res,
            "peaks": {"hbm_GBs": HBM_PEAK, "bf16_tflops": TF_PEAK, "kind": PEAK_KIND}}


def c4(pkg):
    dev = torch.device("cuda:0")
    text = bench.build_shard(pkg, 0, 50_000_000, 384, "bf16", dev)
    image = bench.build_shard(pkg, 0, 10_000_000, 512, "bf16", dev)
    out = []
    for b in (1, 8, 128):
        qt, qi = qdev(b, 384), qdev(b, 512)

        def request():
            t = text.search(qt, 50)
            i = image.search(qi, 12)
            return pkg.fuse(t, i, 4, 0.25)

        ms = timed(request, 10 if b > 1 else 20)
        byts = 50_000_000 * 384 * 2 + 10_000_000 * 512 * 2
        res = request()
        out.append({"batch": b, "ms_per_batch": ms, "fused_requests_per_s": b / (ms * 1e-3),
                    "hbm_GBs": byts / (ms * 1e-3) / 1e9, "hbm_frac": byts / (ms * 1e-3) / 1e9 / HBM_PEAK,
                    "low_conf_fraction": float(res["low_conf"].float().mean().item())})
    text.close(); image.close()
    del text, image
    torch.cuda.empty_cache()
    return {"config": "C4 50M x 384 text (top-50) + 10M x 512 image (top-12) + device fusion + tau gate, rerank off",
            "bytes_per_request_batch": 48_640_000_000, "results": out}


def c5(pkg, cap_rows=140_000_000):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(7)
    sizes = np.exp(rng.uniform(np.log(1_000), np.log(1_000_000), size=1000)).astype(np.int64)
    scale = 1.0
    if sizes.sum() > cap_rows:
        scale = cap_rows / sizes.sum()
        sizes = np.maximum(1_000, (sizes * scale).astype(np.int64))
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    total = int(seg[-1])
    base = bench.build_shard(pkg, 0, total, 512, "bf16", dev)
    ix = pkg.ResidentIndex(base.rows, seg_offsets=seg)
    out = []
    for name, b in (("uniform over tenants", 1000), ("proportional to tenant size", 1000), ("uniform over tenants", 256)):
        p = None if name.startswith("uniform") else sizes / sizes.sum()
        tenants = rng.choice(1000, size=b, p=p).astype(np.int32)
        q = qdev(b, 512)
        ms = timed(lambda: ix.search(q, 10, tenants), 5)
        # K6 groups the queries of a tenant four at a time: a tenant's rows are read once per GROUP.  Algorithmic bytes =
        # rows of the DISTINCT tenants hit (what one ideal pass would read); "rows_read" is what the launch really streams.
        uniq, counts = np.unique(tenants, return_counts=True)
        distinct_rows = int(sizes[uniq].sum())
        rows_read = int((sizes[uniq] * ((counts + 3) // 4)).sum())
        per_query_rows = int(sizes[tenants].sum())
        out.append({"queries": b, "tenant_choice": name, "ms": ms, "distinct_tenants": int(len(uniq)),
                    "algorithmic_rows": distinct_rows, "rows_read_by_the_launch": rows_read,
                    "rows_if_every_query_scanned_alone": per_query_rows,
                    "hbm_GBs_algorithmic": distinct_rows * 1024 / (ms * 1e-3) / 1e9,
                    "hbm_GBs_streamed": rows_read * 1024 / (ms * 1e-3) / 1e9,
                    "hbm_frac_streamed": rows_read * 1024 / (ms * 1e-3) / 1e9 / HBM_PEAK,
                    "qps": b / (ms * 1e-3), "kernel": pkg._native.lib().mmr_last_kernel()})
    ix.close(); base.close()
    return {"config": "C5 1000 tenants, ragged 1k..1M rows x 512 bf16 (log-uniform, seed 7), one varlen launch per batch",
            "total_rows": total, "size_scale_applied": scale, "results": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,4,5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    args = ap.parse_args()
    pkg = importlib.import_module(PKG)
    pkg._native.lib()
    results = {}
    for c, fn in (("1", c1), ("2", c2), ("4", c4), ("5", c5)):
        if c in args.configs.split(","):
            t0 = time.time()
            with bench.ClockSampler(0) as clocks:
                results["C" + c] = fn(pkg)
            results["C" + c]["wall_s"] = time.time() - t0
            results["C" + c]["clocks"] = clocks.summary()
            print(json.dumps({("C" + c): results["C" + c]}), flush=True)
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
