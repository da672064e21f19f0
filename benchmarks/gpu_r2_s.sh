#!/bin/bash
set -u
O=gpurun_out/r2s
mkdir -p $O
for kb in 200 100 64; do
ENC_BENCH_ONLY=cross ENC_BENCH_SKIP_TORCH=1 MMR_ENC_GEMM_SMEM_KB=$kb python benchmarks/encoder_bench.py > $O/enc_cross_smem$kb.json 2>> $O/enc.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s/enc_cross_smem*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], [(r["batch"], r["seq"], round(r["device_encoder_ms"],3)) for r in d["results"]])
    except Exception as e: print(f,"ERR",e)
P
tail -3 $O/enc.err
