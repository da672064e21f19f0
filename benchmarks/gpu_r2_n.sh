#!/bin/bash
set -u
O=gpurun_out/r2n
mkdir -p $O
B=$PWD/multimodal-rag-for-image-text-search_b200/build
python benchmarks/k6_probe.py > $O/k6_probe_items8.json 2> $O/k6_probe.err
for n in 16 24 48; do
MMR_LIB_PATH=$B/libmmr_items$n.so python benchmarks/k6_probe.py > $O/k6_probe_items$n.json 2>> $O/k6_probe.err
MMR_LIB_PATH=$B/libmmr_items$n.so python benchmarks/run_configs.py --configs 5 --out $O/configs_c5_items$n.json > $O/configs$n.log 2>&1
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n/k6_probe_items*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], [(c["case"][:14], round(c["ms"],3), round(c["GBs_streamed"])) for c in d["cases"]])
    except Exception as e: print(f,"ERR",e)
for f in sorted(glob.glob("gpurun_out/r2n/configs_c5_items*.json")):
    try:
        d=json.load(open(f))["C5"]
        print(f.split('/')[-1], [(r["queries"], round(r["ms"],2), round(r["hbm_GBs_streamed"])) for r in d["results"]], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
P
tail -3 $O/k6_probe.err
