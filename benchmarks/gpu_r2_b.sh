#!/bin/bash
# Round-2 GPU call B: new store / mailbox / grouped varlen on one GPU.
set -u
O=gpurun_out/r2b
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/bench_n1.err
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --sweep= > $O/bench_n1_nostore.json 2> $O/bench_n1_nostore.err
tail -3 $O/*.log $O/*.err
