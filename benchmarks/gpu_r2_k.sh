#!/bin/bash
set -u
O=gpurun_out/r2k
mkdir -p $O
for B in 8 128; do
python benchmarks/k2_probe.py --rows 1000000 --batch $B --reps 200 "MMR_UMMA_FUSED_PROBE=0" > $O/k2_1m_b$B.json 2> $O/k2_1m_b$B.err
done
python benchmarks/k2_probe.py --rows 10000000 --batch 8 --reps 50 "MMR_UMMA_FUSED_PROBE=0" > $O/k2_10m_b8.json 2>> $O/k2_1m_b8.err
cat $O/k2_*.json | grep -E '"switches"|"ms"' | paste - -
