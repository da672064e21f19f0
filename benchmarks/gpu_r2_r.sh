#!/bin/bash
set -u
O=gpurun_out/r2r
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_encoders.py -m gpu -q > $O/pytest_enc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_enc.log
MMR_ENC_ATT_MMA=0 timeout 900 python -m pytest tests/test_gpu_encoders.py -m gpu -q > $O/pytest_enc_nofuse.log 2>&1; echo "pytest rc=$?" >> $O/pytest_enc_nofuse.log
python benchmarks/encoder_bench.py > $O/encoder_bench.json 2> $O/encoder_bench.err
MMR_ENC_ATT_MMA=0 python benchmarks/encoder_bench.py > $O/encoder_bench_nofuse.json 2>> $O/encoder_bench.err
grep -E "^E  |passed|failed|Error" $O/pytest_enc.log | head -30; grep -E "^E  |passed|failed|Error" $O/pytest_enc_nofuse.log | head -20
python - <<'P'
import json
for f in ("encoder_bench.json","encoder_bench_nofuse.json"):
    try:
        d=json.load(open("gpurun_out/r2r/"+f)); print(f, [(r["model"], r["batch"], r["seq"], round(r["device_encoder_ms"],3), r["kernel_launches"]) for r in d["results"]])
    except Exception as e: print(f,"ERR",e)
P
tail -3 $O/encoder_bench.err
