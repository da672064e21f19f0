#!/bin/bash
set -u
O=gpurun_out/r2l
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_fuzz.py tests/test_gpu_round2.py tests/test_gpu_configs_at_size.py tests/test_gpu_fuse_store.py -m gpu -q > $O/pytest_scan.log 2>&1; echo "pytest rc=$?" >> $O/pytest_scan.log
python benchmarks/run_configs.py --configs 5 --out $O/configs_c5.json > $O/configs.log 2>&1; echo "rc=$?" >> $O/configs.log
tail -n 4 $O/pytest_scan.log
