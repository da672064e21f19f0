#!/bin/bash
set -u
O=gpurun_out/r2y
mkdir -p $O
S="MMR_UMMA_STAGES=2 MMR_UMMA_STAGES=3 MMR_UMMA_STAGES=4 MMR_UMMA_MODE=ts MMR_UMMA_MODE=ts,MMR_UMMA_STAGES=4 MMR_UMMA_MODE=ts,MMR_UMMA_STAGES=8"
python benchmarks/k2_probe.py --rows 10000000 --batch 8 --reps 30 $S > $O/k2_10m_b8.json 2> $O/err.log
python benchmarks/k2_probe.py --rows 10000000 --batch 128 --reps 30 $S > $O/k2_10m_b128.json 2>> $O/err.log
python benchmarks/k2_probe.py --rows 1000000 --batch 8 --reps 200 $S > $O/k2_1m_b8.json 2>> $O/err.log
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2y/k2_*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], [(r["switches"].replace("MMR_UMMA_",""), round(r["ms"],4)) for r in d["results"]])
    except Exception as e: print(f,"ERR",e)
P
tail -3 $O/err.log
