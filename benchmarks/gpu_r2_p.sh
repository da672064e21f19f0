#!/bin/bash
set -u
O=gpurun_out/r2p
mkdir -p $O
# ncu --set full of the varlen kernel: NQ = 1 items (64 equal tenants) and NQ = 4 items (4 + 4 queries)
python benchmarks/k6_probe.py --reps 3 --only "64 queries" > $O/plain1.json 2> $O/plain1.err && \
ncu --set full --clock-control none --import-source on -k regex:scan_stream -s 12 -c 1 -o $O/k6_nq1 python benchmarks/k6_probe.py --reps 3 --only "64 queries" > $O/ncu1.log 2>&1
python benchmarks/k6_probe.py --reps 3 --only "4 + 4" > $O/plain4.json 2> $O/plain4.err && \
ncu --set full --clock-control none --import-source on -k regex:scan_stream -s 12 -c 1 -o $O/k6_nq4 python benchmarks/k6_probe.py --reps 3 --only "4 + 4" > $O/ncu4.log 2>&1
for n in k6_nq1 k6_nq4; do
  ncu -i $O/$n.ncu-rep --page raw --csv > $O/${n}_raw.csv 2>/dev/null
  ncu -i $O/$n.ncu-rep --page source --csv --print-source sass > $O/${n}_source.csv 2>/dev/null
  ncu -i $O/$n.ncu-rep --page details --csv > $O/${n}_details.csv 2>/dev/null
done
ls -la $O
rm -f $O/*.ncu-rep
tail -5 $O/ncu4.log
