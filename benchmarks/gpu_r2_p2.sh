#!/bin/bash
set -u
O=gpurun_out/r2p2
mkdir -p $O
python benchmarks/k6_probe.py --reps 3 --only "4 + 4" > $O/plain4.json 2> $O/plain4.err && \
ncu --set full --clock-control none -k regex:scan_stream -s 12 -c 1 -o $O/k6_nq4 python benchmarks/k6_probe.py --reps 3 --only "4 + 4" > $O/ncu4.log 2>&1
ncu -i $O/k6_nq4.ncu-rep --page raw --csv > $O/k6_nq4_12warps_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls -la $O; tail -2 $O/ncu4.log
