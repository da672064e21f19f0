#!/bin/bash
# same-box A/B: library as of the start of this session (637fa21) vs the current one, K2 at B = 8 / 128 / 1024 and K1
set -u
O=gpurun_out/r2ab
mkdir -p $O
B=$PWD/multimodal-rag-for-image-text-search_b200/build
for rep in 1 2; do
for lib in cur old; do
  if [ $lib = old ]; then export MMR_LIB_PATH=$B/libmmr_r2start.so; else unset MMR_LIB_PATH; fi
  python benchmarks/k2_probe.py --rows 10000000 --batch 1024 --reps 20 > $O/k2_b1024_${lib}_$rep.json 2>> $O/err.log
  python benchmarks/k2_probe.py --rows 10000000 --batch 128 --reps 30 > $O/k2_b128_${lib}_$rep.json 2>> $O/err.log
  python benchmarks/k2_probe.py --rows 10000000 --batch 8 --reps 30 > $O/k2_b8_${lib}_$rep.json 2>> $O/err.log
  python benchmarks/k2_probe.py --rows 1000000 --batch 1024 --reps 100 > $O/k2_1m_b1024_${lib}_$rep.json 2>> $O/err.log
done
done
unset MMR_LIB_PATH
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2ab/k2_*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], [round(r["ms"],4) for r in d["results"]])
    except Exception as e: print(f,"ERR",e)
P
nvidia-smi --query-gpu=clocks.sm,power.draw,power.limit,temperature.gpu --format=csv
tail -3 $O/err.log
