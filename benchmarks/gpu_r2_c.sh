#!/bin/bash
# Round-2 GPU call C (N GPUs): full GPU suite incl. the multi-GPU tests, SPMD bench, single-process multi-GPU store bench.
set -u
N=${1:-2}
O=gpurun_out/r2c_n$N
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?" >> $O/bench_n$N.err
python benchmarks/multi_store_bench.py --gpus $N > $O/multi_store_n$N.json 2> $O/multi_store_n$N.err; echo "rc=$?" >> $O/multi_store_n$N.err
tail -4 $O/pytest_gpu.log $O/*.err
