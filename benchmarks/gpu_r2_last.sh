#!/bin/bash
set -u
O=gpurun_out/r2last
mkdir -p $O
python benchmarks/encoder_bench.py > $O/encoder_bench.json 2> $O/encoder_bench.err
python benchmarks/run_configs.py --configs 1,2 --out $O/configs_c1_c2.json > $O/configs.log 2>&1
python - <<'P'
import json
e=json.load(open("gpurun_out/r2last/encoder_bench.json"))
print([(r["model"], r["batch"], r["seq"], round(r["device_encoder_ms"],3), round(r["torch_fp32_ms"],2)) for r in e["results"]])
d=json.load(open("gpurun_out/r2last/configs_c1_c2.json"))
print({k:round(v*1000,1) for k,v in d["C1"].items() if k.startswith("ms_")})
print([(r["batch"], round(r["ms"],4)) for r in d["C2"]["sweep"]])
P
