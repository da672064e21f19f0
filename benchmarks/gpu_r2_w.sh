#!/bin/bash
set -u
O=gpurun_out/r2w2
mkdir -p $O
B=$PWD/multimodal-rag-for-image-text-search_b200/build
for rep in 1 2; do
ENC_BENCH_ONLY=cross ENC_BENCH_SKIP_TORCH=1 python benchmarks/encoder_bench.py > $O/enc_cross_libm_$rep.json 2>> $O/enc.err
MMR_LIB_PATH=$B/libmmr_fasterf.so ENC_BENCH_ONLY=cross ENC_BENCH_SKIP_TORCH=1 python benchmarks/encoder_bench.py > $O/enc_cross_fast_$rep.json 2>> $O/enc.err
done
MMR_LIB_PATH=$B/libmmr_fasterf.so timeout 600 python -m pytest tests/test_gpu_encoders.py -m gpu -q > $O/pytest_enc_fast.log 2>&1; echo "pytest rc=$?" >> $O/pytest_enc_fast.log
grep -E "^E  |passed|failed|Error|rc=" $O/pytest_enc_fast.log | head
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2w2/enc_*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], [(r["batch"], r["seq"], round(r["device_encoder_ms"],3)) for r in d["results"]])
    except Exception as e: print(f,"ERR",e)
P
tail -3 $O/enc.err
