#!/bin/bash
set -u
O=gpurun_out/r2e
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python benchmarks/fixed_cost.py > $O/fixed_cost.json 2> $O/fixed_cost.err; echo "rc=$?" >> $O/fixed_cost.err
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/bench_n1.err
python benchmarks/run_configs.py --configs 1,2 --out $O/configs_c1_c2.json > $O/configs.log 2>&1; echo "rc=$?" >> $O/configs.log
tail -n 5 $O/*.log $O/*.err
