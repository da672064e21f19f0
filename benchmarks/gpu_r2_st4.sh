#!/bin/bash
set -u
O=gpurun_out/r2st4
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fuse_store.py tests/test_gpu_round2.py tests/test_gpu_c_abi.py tests/test_gpu_configs_at_size.py -m gpu -q > $O/pytest_store.log 2>&1; echo "pytest rc=$?" >> $O/pytest_store.log
grep -E "^E  |passed|failed|Error|rc=" $O/pytest_store.log | head
python benchmarks/store_overhead.py > $O/ov_new.json 2>> $O/err.log; cat $O/ov_new.json
python bench.py --steps 100 --warmup 10 --sweep= > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
python - <<'P'
import json
b=json.loads(open("gpurun_out/r2st4/bench_n1.json").read().strip().splitlines()[-1]); print(b["value"], b["ms_per_step"], b["e2e"]["value"], b["e2e"]["ms_per_step"], b["e2e"].get("c_abi_ms_per_step"))
P
tail -2 $O/err.log $O/bench_n1.err
