#!/bin/bash
set -u
O=gpurun_out/r2st
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fuse_store.py tests/test_gpu_round2.py tests/test_gpu_c_abi.py -m gpu -q > $O/pytest_store.log 2>&1; echo "pytest rc=$?" >> $O/pytest_store.log
python benchmarks/run_configs.py --configs 1 --out $O/configs_c1.json > $O/configs.log 2>&1
python bench.py --steps 100 --warmup 10 --sweep= > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
grep -E "^E  |passed|failed|Error|rc=" $O/pytest_store.log | head
python - <<'P'
import json
d=json.load(open("gpurun_out/r2st/configs_c1.json"))["C1"]; print({k:v for k,v in d.items() if k.startswith("ms_")})
b=json.loads(open("gpurun_out/r2st/bench_n1.json").read().strip().splitlines()[-1]); print(b["value"], b["ms_per_step"], b["e2e"]["value"], b["e2e"]["ms_per_step"], b["e2e"].get("c_abi_ms_per_step"))
P
tail -2 $O/bench_n1.err
