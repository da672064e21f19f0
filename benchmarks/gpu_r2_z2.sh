#!/bin/bash
set -u
O=gpurun_out/r2z2
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_encoders.py -m gpu -q > $O/pytest_enc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_enc.log
ENC_BENCH_ONLY=cross ENC_BENCH_SKIP_TORCH=1 python benchmarks/encoder_bench.py > $O/enc_cross.json 2>> $O/enc.err
ENC_BENCH_ONLY=clip_text ENC_BENCH_SKIP_TORCH=1 python benchmarks/encoder_bench.py > $O/enc_clip.json 2>> $O/enc.err
grep -E "^E  |passed|failed|Error|rc=" $O/pytest_enc.log | head
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z2/enc_*.json")):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], [(r["batch"], r["seq"], round(r["device_encoder_ms"],3)) for r in d["results"]])
    except Exception as e: print(f,"ERR",e)
P
tail -3 $O/enc.err
