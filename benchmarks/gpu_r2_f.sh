#!/bin/bash
set -u
O=gpurun_out/r2f
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
K2="python bench.py --batch 1024 --steps 20 --warmup 5 --no-cpu-baseline --sweep="
MMR_UMMA_LOCKSTEP=0 $K2 > $O/k2_b1024_lockstep0.json 2> $O/k2_l0.err
MMR_UMMA_LOCKSTEP=1 $K2 > $O/k2_b1024_lockstep1.json 2> $O/k2_l1.err
MMR_UMMA_LOCKSTEP=0 $K2 > $O/k2_b1024_lockstep0b.json 2>> $O/k2_l0.err
MMR_UMMA_LOCKSTEP=1 $K2 > $O/k2_b1024_lockstep1b.json 2>> $O/k2_l1.err
K2S="python bench.py --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --sweep="
$K2S > $O/plain_k2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_umma2 -s 7 -c 1 -f -o $O/r02_k2pair_lockstep_b1024 $K2S > $O/ncu_k2.log 2>&1
python benchmarks/encoder_bench.py > $O/encoder_bench.json 2> $O/encoder_bench.err; echo "rc=$?" >> $O/encoder_bench.err
python benchmarks/fixed_cost.py > $O/fixed_cost.json 2> $O/fixed_cost.err
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/bench_n1.err
tail -n 5 $O/pytest_gpu.log $O/*.err
