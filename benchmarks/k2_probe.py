#!/usr/bin/env python
"""K2 experiments at one batch size: ms per search under different library switches (measurement only).

    python benchmarks/k2_probe.py [--batch 1024] [--rows 10000000] "MMR_UMMA_SKIP_EPI=1" "MMR_UMMA_PAIR=0" ...

Each positional argument is a comma-separated list of NAME=VALUE switches applied for one timing; the first timing is
always the default configuration."""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("settings", nargs="*")
    args = ap.parse_args()
    pkg = importlib.import_module("multimodal-rag-for-image-text-search_b200")
    N = pkg._native
    ix = bench.build_shard(pkg, 0, args.rows, args.dim, "bf16", torch.device("cuda:0"))
    q = torch.from_numpy(bench.gen_queries(args.batch, args.dim)).cuda()
    out = []
    for setting in [""] + list(args.settings):
        pairs = [kv.split("=") for kv in setting.split(",") if kv]
        for name, value in pairs:
            N.set_option(name, value)
        for _ in range(3):
            ix.search(q, args.k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            ix.search(q, args.k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        tfl = 2.0 * args.batch * args.rows * args.dim / (ms * 1e-3) / 1e12
        out.append({"switches": setting or "(default)", "ms": ms, "tflops": tfl, "frac_of_1639.9": tfl / 1639.9})
        for name, _ in pairs:
            N.set_option(name, None)
    print(json.dumps({"batch": args.batch, "rows": args.rows, "results": out}, indent=1))


if __name__ == "__main__":
    main()
