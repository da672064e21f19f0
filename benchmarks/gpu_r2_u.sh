#!/bin/bash
set -u
O=gpurun_out/r2u
mkdir -p $O
for shape in "128 128" "8 512"; do
set -- $shape
python benchmarks/encoder_profile.py cross $1 $2 > $O/plain_$1x$2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$1x$2.csv python benchmarks/encoder_profile.py cross $1 $2 > $O/ncu_$1x$2.log 2>&1
python benchmarks/ncu_excerpt.py list $O/launches_$1x$2.csv $O/r02_encoder_cross_b$1_s$2_launches.csv > $O/excerpt_$1x$2.log 2>&1
tail -n 14 $O/r02_encoder_cross_b$1_s$2_launches.csv
done
