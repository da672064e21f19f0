#!/bin/bash
set -u
O=gpurun_out/r2g
mkdir -p $O
for cfg in "cross 8 128" "minilm 1 16" "clip 1 16" "cross 8 512"; do
  tag=$(echo $cfg | tr ' ' '_')
  python benchmarks/encoder_profile.py $cfg > $O/plain_$tag.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/enc_$tag.csv python benchmarks/encoder_profile.py $cfg > $O/ncu_$tag.log 2>&1
done
ls -la $O
