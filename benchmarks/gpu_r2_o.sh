#!/bin/bash
set -u
O=gpurun_out/r2o
mkdir -p $O
B=$PWD/multimodal-rag-for-image-text-search_b200/build
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_fuzz.py tests/test_gpu_round2.py tests/test_gpu_configs_at_size.py -m gpu -q -x > $O/pytest_scan.log 2>&1; echo "pytest rc=$?" >> $O/pytest_scan.log
for n in s22 s33 s34 s24; do
MMR_LIB_PATH=$B/libmmr_$n.so python benchmarks/k6_probe.py > $O/k6_probe_$n.json 2>> $O/k6_probe.err
MMR_LIB_PATH=$B/libmmr_$n.so python benchmarks/run_configs.py --configs 5 --out $O/configs_c5_$n.json > $O/configs_$n.log 2>&1
done
MMR_LIB_PATH=$B/libmmr_s33.so python benchmarks/run_configs.py --configs 2 --out $O/configs_c2_s33.json > $O/configs_c2_s33.log 2>&1
MMR_LIB_PATH=$B/libmmr_s22.so python benchmarks/run_configs.py --configs 2 --out $O/configs_c2_s22.json > $O/configs_c2_s22.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o/k6_probe_s*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], [(c["case"][:12], round(c["ms"],3), round(c["GBs_streamed"])) for c in d["cases"]])
    except Exception as e: print(f,"ERR",e)
for f in sorted(glob.glob("gpurun_out/r2o/configs_c5_s*.json")):
    try:
        d=json.load(open(f))["C5"]
        print(f.split('/')[-1], [(r["queries"], round(r["ms"],2), round(r["hbm_GBs_streamed"])) for r in d["results"]], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
for f in sorted(glob.glob("gpurun_out/r2o/configs_c2_s*.json")):
    try:
        d=json.load(open(f))["C2"]
        print(f.split('/')[-1], [(r["batch"], round(r["ms"],4)) for r in d["sweep"]])
    except Exception as e: print(f, "ERR", e)
P
tail -3 $O/pytest_scan.log; tail -3 $O/k6_probe.err
