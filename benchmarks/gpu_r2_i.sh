#!/bin/bash
set -u
O=gpurun_out/r2i
mkdir -p $O
for B in 8 128; do
python benchmarks/k2_probe.py --rows 1000000 --batch $B --reps 200 "MMR_UMMA_NOPROBE=1" "MMR_UMMA_MODE=ts" "MMR_UMMA_MODE=ts,MMR_UMMA_NOPROBE=1" > $O/k2_1m_b$B.json 2> $O/k2_1m_b$B.err
done
python benchmarks/k2_probe.py --rows 2500000 --batch 8 --reps 100 "MMR_UMMA_NOPROBE=1" > $O/k2_2p5m_b8.json 2>> $O/k2_1m_b8.err
python benchmarks/k2_probe.py --rows 5000000 --batch 8 --reps 100 "MMR_UMMA_NOPROBE=1" > $O/k2_5m_b8.json 2>> $O/k2_1m_b8.err
cat $O/*.json | grep -E '"switches"|"ms"'
