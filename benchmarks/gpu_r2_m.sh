#!/bin/bash
set -u
O=gpurun_out/r2m
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_fuzz.py tests/test_gpu_round2.py tests/test_gpu_configs_at_size.py tests/test_gpu_fuse_store.py -m gpu -q -x > $O/pytest_scan.log 2>&1; echo "pytest rc=$?" >> $O/pytest_scan.log
python benchmarks/run_configs.py --configs 5 --out $O/configs_c5.json > $O/configs.log 2>&1; echo "rc=$?" >> $O/configs.log
MMR_LIB_PATH=$PWD/multimodal-rag-for-image-text-search_b200/build/libmmr_items24.so python benchmarks/run_configs.py --configs 5 --out $O/configs_c5_items24.json > $O/configs24.log 2>&1; echo "rc=$?" >> $O/configs24.log
tail -n 4 $O/pytest_scan.log
python - <<'P'
import json
for f in ("configs_c5.json","configs_c5_items24.json"):
    try:
        d=json.load(open("gpurun_out/r2m/"+f))["C5"]
        print(f, [(r["queries"], round(r["ms"],2), round(r["hbm_GBs_streamed"])) for r in d["results"]], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
P
