#!/bin/bash
# Round-2 GPU call A: state check of the round-1 kernels + ncu captures of the SHIPPED K1 ring and K2 pair mode.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.csv 2>&1
free -g > $O/free.txt; nproc >> $O/free.txt
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
K1CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sweep="
$K1CMD > $O/plain_k1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_stream -s 3 -c 2 -f -o $O/r02_k1_full $K1CMD > $O/ncu_k1.log 2>&1
$K1CMD > $O/plain_k1b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_k1_launches.csv $K1CMD > $O/ncu_k1l.log 2>&1
K2CMD="python bench.py --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --sweep="
$K2CMD > $O/plain_k2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_umma2 -s 7 -c 1 -f -o $O/r02_k2pair_b1024 $K2CMD > $O/ncu_k2.log 2>&1
$K2CMD > $O/plain_k2b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_k2_launches.csv $K2CMD > $O/ncu_k2l.log 2>&1
ls -la $O
