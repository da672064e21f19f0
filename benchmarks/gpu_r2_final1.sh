#!/bin/bash
# Round-2 final single-GPU evidence run.
set -u
O=gpurun_out/r2final
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.csv 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" >> $O/bench_n1.err
python bench.py --steps 200 --warmup 20 > $O/bench_n1_long.json 2> $O/bench_n1_long.err
python benchmarks/run_configs.py --configs 1,2,4,5 --out $O/configs.json > $O/configs.log 2>&1; echo "rc=$?" >> $O/configs.log
python benchmarks/fixed_cost.py > $O/fixed_cost.json 2> $O/fixed_cost.err
python benchmarks/encoder_bench.py > $O/encoder_bench.json 2> $O/encoder_bench.err
# ncu: K1 (shipped), K2 pair B=1024, encoder GEMM + attention
K1CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sweep="
$K1CMD > $O/plain_k1.log 2>&1 && \
ncu --set full --clock-control none -k regex:scan_stream -s 3 -c 2 -f -o $O/r02_k1_final $K1CMD > $O/ncu_k1.log 2>&1
ncu -i $O/r02_k1_final.ncu-rep --page raw --csv > $O/r02_k1_final_raw.csv 2>/dev/null; rm -f $O/r02_k1_final.ncu-rep
K2CMD="python bench.py --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --sweep="
$K2CMD > $O/plain_k2.log 2>&1 && \
ncu --set full --clock-control none -k regex:scan_umma2 -s 7 -c 1 -f -o $O/r02_k2pair_final $K2CMD > $O/ncu_k2.log 2>&1
ncu -i $O/r02_k2pair_final.ncu-rep --page raw --csv > $O/r02_k2pair_final_raw.csv 2>/dev/null; rm -f $O/r02_k2pair_final.ncu-rep
ECMD="python benchmarks/encoder_profile.py cross 64 128"
$ECMD > $O/plain_enc.log 2>&1 && \
ncu --set full --clock-control none -k regex:"gemm_wt_kernel|attention_kernel" -s 24 -c 6 -f -o $O/r02_encoder $ECMD > $O/ncu_enc.log 2>&1
ncu -i $O/r02_encoder.ncu-rep --page raw --csv > $O/r02_encoder_raw.csv 2>/dev/null; rm -f $O/r02_encoder.ncu-rep
$ECMD > $O/plain_enc2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_encoder_launches.csv $ECMD > $O/ncu_encl.log 2>&1
python benchmarks/k2_probe.py --batch 1024 "MMR_UMMA_SKIP_EPI=1" "MMR_UMMA_PAIR=0" "MMR_UMMA_PAIR=0,MMR_UMMA_MODE=ss" "MMR_UMMA_NOPROBE=1" > $O/k2_probe_b1024.json 2> $O/k2_probe.err
python benchmarks/k2_probe.py --batch 4096 --reps 6 "MMR_UMMA_SKIP_EPI=1" > $O/k2_probe_b4096.json 2>> $O/k2_probe.err
du -sh $O; ls -la $O | head -60
tail -n 4 $O/smoke.log $O/pytest_gpu.log $O/configs.log
