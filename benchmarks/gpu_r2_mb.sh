#!/bin/bash
set -u
O=gpurun_out/r2mb
mkdir -p $O
for rep in 1 2; do
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --sweep= > $O/bench_sync_$rep.json 2>> $O/err.log
MMR_MAILBOX=1 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --sweep= > $O/bench_mailbox_$rep.json 2>> $O/err.log
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2mb/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],4), "cabi", round(d["e2e"]["c_abi_ms_per_step"],4))
    except Exception as e: print(f,"ERR",e)
P
tail -3 $O/err.log
