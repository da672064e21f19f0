#!/bin/bash
# 2 GPUs: multi-GPU tests + SPMD bench at N=2 after the round's kernel changes
set -u
O=gpurun_out/r2x
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
python benchmarks/multi_store_bench.py --gpus 2 > $O/multi_store_n2.json 2> $O/multi_store_n2.err; echo "rc=$?" >> $O/multi_store_n2.err
grep -E "passed|failed|rc=" $O/pytest_multi.log; tail -n 2 $O/bench_n2.err $O/multi_store_n2.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2x/bench_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","parity_checked","n_gpus")}, d["e2e"]["value"])
P
