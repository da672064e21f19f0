#!/bin/bash
# Round-2 GPU call D: encoders + everything else on one GPU, fixed-cost breakdown.
set -u
O=gpurun_out/r2d
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_encoders.py -q -x > $O/pytest_enc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_enc.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_encoders.py > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python benchmarks/fixed_cost.py > $O/fixed_cost.json 2> $O/fixed_cost.err; echo "rc=$?" >> $O/fixed_cost.err
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/bench_n1.err
tail -n 5 $O/*.log $O/*.err
