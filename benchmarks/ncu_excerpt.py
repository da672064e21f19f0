#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small, tracked files under profiles/.

    python benchmarks/ncu_excerpt.py raw  <in.ncu-rep> <out.csv>      # transposed `--page raw` (metric, unit, launch0..)
    python benchmarks/ncu_excerpt.py list <launches.csv> <out.csv>    # id, kernel (short), grid, block, ns  + share per kernel
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def raw(src, dst):
    text = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        for i, h in enumerate(hdr):
            if h in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
                continue
            w.writerow([h, units[i]] + [r[i] for r in launches])


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name if len(name) < 90 else name[:87] + "..."


def launch_list(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, gi, bi, vi, ii = (hdr.index(x) for x in ("Kernel Name", "Grid Size", "Block Size", "Metric Value", "ID"))
    total = defaultdict(lambda: [0, 0.0])
    body = []
    for r in rows[1:]:
        k = short(r[ki])
        ns = float(r[vi].replace(",", ""))
        total[k][0] += 1
        total[k][1] += ns
        body.append([r[ii], k, r[gi], r[bi], int(ns)])
    grand = sum(v[1] for v in total.values())
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["# per-kernel totals: kernel, launches, total_ns, share_of_all_launches"])
        for k, (n, ns) in sorted(total.items(), key=lambda kv: -kv[1][1]):
            w.writerow(["#", k, n, int(ns), f"{ns / grand:.4f}"])
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum_ns"])
        w.writerows(body)


if __name__ == "__main__":
    {"raw": raw, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
