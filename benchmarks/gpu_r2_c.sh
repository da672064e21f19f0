#!/bin/bash
# Round-2 GPU call C (N GPUs): the multi-GPU tests, SPMD bench at N (and N/2), single-process multi-GPU store bench.
set -u
N=${1:-2}
O=gpurun_out/r2c_n$N
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -v > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
for G in $N $((N/2)); do
  [ $G -ge 2 ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $G --steps 200 --warmup 20 > $O/bench_n$G.json 2> $O/bench_n$G.err; echo "rc=$?" >> $O/bench_n$G.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 --batch 128 > $O/bench_n${N}_b128.json 2> $O/bench_n${N}_b128.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 10 --batch 1024 > $O/bench_n${N}_b1024.json 2> $O/bench_n${N}_b1024.err
python benchmarks/multi_store_bench.py --gpus $N > $O/multi_store_n$N.json 2> $O/multi_store_n$N.err; echo "rc=$?" >> $O/multi_store_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 10 --warmup 3 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err
tail -n 6 $O/pytest_multi.log $O/*.err
