#!/usr/bin/env python
"""Fixed cost of one search: latency of the host-buffer call and of the device-resident call vs table size.

    python benchmarks/fixed_cost.py > gpurun_out/fixed_cost.json

For rows in {4k .. 4M} x 512 bf16, batch 1, top-10: median microseconds of (a) mmr_search_host with the query in the kernel
parameters + mailbox, (b) the same with an H2D query copy, (c) with a stream synchronisation instead of the mailbox, (d) the
device-resident call timed with CUDA events back to back (pipelined launches on/off).  The intercept of (a) over rows is
the per-request fixed cost a latency-bound deployment pays."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pkg = importlib.import_module("multimodal-rag-for-image-text-search_b200")
N = pkg._native


def host_us(ix, q, reps=300):
    for i in range(20):
        ix.search_host(q[i % len(q)], 10)
    ts = []
    for i in range(reps):
        t0 = time.perf_counter()
        ix.search_host(q[i % len(q)], 10)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e6)


def device_us(ix, qd, reps=300):
    out = (torch.empty((1, 10), dtype=torch.float32, device="cuda"), torch.empty((1, 10), dtype=torch.int64, device="cuda"))
    for i in range(20):
        ix.search(qd[i % len(qd)], 10, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ix.search(qd[i % len(qd)], 10, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    q = bench.gen_queries(64, 512)
    qd = torch.from_numpy(q).cuda()[:, None, :].contiguous()
    rows_all = [4_096, 65_536, 262_144, 1_048_576, 4_194_304]
    big = bench.build_shard(pkg, 0, rows_all[-1], 512, "bf16", torch.device("cuda:0"))
    out = []
    for n in rows_all:
        ix = pkg.ResidentIndex(big.rows[:n])
        rec = {"rows": n, "ideal_us_at_7.4TBs": n * 1024 / 7.4e12 * 1e6}
        for name, inline, mailbox in (("host_inline_mailbox", "1", "1"), ("host_h2d_mailbox", "0", "1"), ("host_inline_sync", "1", "0"),
                                      ("host_h2d_sync", "0", "0")):
            N.set_option("MMR_INLINE_QUERY", inline)
            N.set_option("MMR_MAILBOX", mailbox)
            rec[name + "_us"] = host_us(ix, q[:, None, :])
        N.set_option("MMR_INLINE_QUERY", None)
        N.set_option("MMR_MAILBOX", None)
        for pdl in ("0", "1"):
            N.set_option("MMR_PDL", pdl)
            rec[f"device_pdl{pdl}_us"] = device_us(ix, qd)
        N.set_option("MMR_PDL", None)
        out.append(rec)
        ix.close()
    print(json.dumps({"what": "batch 1, top-10, 512-d bf16; median host latency / mean device time per search", "results": out}, indent=1))


if __name__ == "__main__":
    main()
