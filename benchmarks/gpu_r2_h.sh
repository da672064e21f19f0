#!/bin/bash
set -u
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_encoders.py -q > $O/pytest_enc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_enc.log
python benchmarks/encoder_bench.py > $O/encoder_bench.json 2> $O/encoder_bench.err; echo "rc=$?" >> $O/encoder_bench.err
for cfg in "cross 8 128" "minilm 1 16"; do
  tag=$(echo $cfg | tr ' ' '_')
  python benchmarks/encoder_profile.py $cfg > $O/plain_$tag.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/enc_$tag.csv python benchmarks/encoder_profile.py $cfg > $O/ncu_$tag.log 2>&1
done
tail -n 4 $O/pytest_enc.log $O/*.err
