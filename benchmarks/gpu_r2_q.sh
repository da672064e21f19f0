#!/bin/bash
set -u
O=gpurun_out/r2q
mkdir -p $O
B=$PWD/multimodal-rag-for-image-text-search_b200/build
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_fuzz.py tests/test_gpu_round2.py tests/test_gpu_configs_at_size.py tests/test_gpu_fuse_store.py -m gpu -q -x > $O/pytest_scan.log 2>&1; echo "pytest rc=$?" >> $O/pytest_scan.log
for n in nw12 nw8; do
MMR_LIB_PATH=$B/libmmr_$n.so python benchmarks/k6_probe.py > $O/k6_probe_$n.json 2>> $O/k6_probe.err
MMR_LIB_PATH=$B/libmmr_$n.so python benchmarks/run_configs.py --configs 5 --out $O/configs_c5_$n.json > $O/configs_$n.log 2>&1
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2q/k6_probe_*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], [(c["case"][:12], round(c["ms"],3), round(c["GBs_streamed"])) for c in d["cases"]])
    except Exception as e: print(f,"ERR",e)
for f in sorted(glob.glob("gpurun_out/r2q/configs_c5_*.json")):
    try:
        d=json.load(open(f))["C5"]
        print(f.split('/')[-1], [(r["queries"], round(r["ms"],2), round(r["hbm_GBs_streamed"])) for r in d["results"]], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
P
tail -3 $O/pytest_scan.log; tail -3 $O/k6_probe.err
