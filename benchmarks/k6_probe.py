"""K6 (grouped varlen launch) next to K1 on the same rows: what the work-item machinery itself costs.

  python benchmarks/k6_probe.py [--rows 10000000]

Cases on one resident table of `rows` x 512 bf16:
  * K1 uniform, 1 query, whole table                         (the roofline case: 10.24 GB per launch)
  * K6, 2 queries on the 2 halves of the table               (same bytes, ~148 x items-per-CTA equal pieces: no imbalance)
  * K6, 64 queries on 64 equal tenants                       (same bytes, many items)
  * K6, 8 queries = 4 + 4 on the 2 halves                    (NQ = 4 class only: FMA-paced arithmetic)
Prints one JSON object; CUDA events, 3 warm-up + `reps` timed launches per case.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

PKG = "multimodal-rag-for-image-text-search_b200"


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="", help="substring: run only the K6 cases whose name contains it (for ncu)")
    a = ap.parse_args()
    pkg = importlib.import_module(PKG)
    lib = pkg._native.lib()
    dev = torch.device("cuda:0")
    base = bench.build_shard(pkg, 0, a.rows, 512, "bf16", dev)
    peak = bench.measured_peaks()[0]
    gb = a.rows * 1024 / 1e9
    out = {"rows": a.rows, "gb_per_pass": gb, "hbm_peak": peak, "cases": []}

    def case(name, ntenants, tenants, passes=1.0):
        if a.only and a.only not in name:
            return
        seg = np.linspace(0, a.rows, ntenants + 1).astype(np.int64)
        ix = pkg.ResidentIndex(base.rows, seg_offsets=seg)
        q = torch.from_numpy(bench.gen_queries(len(tenants), 512)).cuda()
        t = None if tenants is None else np.asarray(tenants, dtype=np.int32)
        ms = timed(lambda: ix.search(q, 10, t), a.reps)
        out["cases"].append({"case": name, "ms": ms, "kernel": int(lib.mmr_last_kernel()), "GBs_streamed": gb * passes / (ms * 1e-3),
                             "frac": gb * passes / (ms * 1e-3) / peak})
        ix.close()

    ix1 = pkg.ResidentIndex(base.rows)
    q1 = torch.from_numpy(bench.gen_queries(1, 512)).cuda()
    ms = timed(lambda: ix1.search(q1, 10), a.reps)
    out["cases"].append({"case": "K1 uniform, 1 query", "ms": ms, "kernel": int(lib.mmr_last_kernel()), "GBs_streamed": gb / (ms * 1e-3),
                         "frac": gb / (ms * 1e-3) / peak})
    q2 = torch.from_numpy(bench.gen_queries(2, 512)).cuda()
    ms = timed(lambda: ix1.search(q2, 10), a.reps)
    out["cases"].append({"case": "K1 uniform, 2 queries (NQ = 2)", "ms": ms, "kernel": int(lib.mmr_last_kernel()), "GBs_streamed": gb / (ms * 1e-3),
                         "frac": gb / (ms * 1e-3) / peak})
    ix1.close()
    case("K6, 2 queries on 2 half-table tenants", 2, [0, 1])
    case("K6, 64 queries on 64 equal tenants", 64, list(range(64)))
    case("K6, 1000 queries on 1000 equal tenants", 1000, list(range(1000)))
    case("K6, 2 + 2 queries on 2 half-table tenants (NQ = 2 items)", 2, [0, 0, 1, 1])
    case("K6, 4 + 4 queries on 2 half-table tenants (NQ = 4 items)", 2, [0, 0, 0, 0, 1, 1, 1, 1])
    base.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
