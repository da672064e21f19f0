#!/bin/bash
set -u
O=gpurun_out/r2j
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python benchmarks/encoder_bench.py > $O/encoder_bench.json 2> $O/encoder_bench.err
python benchmarks/run_configs.py --configs 2 --out $O/configs_c2.json > $O/configs.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
tail -n 4 $O/pytest_gpu.log
