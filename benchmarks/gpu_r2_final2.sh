#!/bin/bash
# Round-2 final single-GPU evidence run (after the K6 / encoder changes of the second session).
set -u
O=gpurun_out/r2final2
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.csv 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err; echo "bench rc=$?" >> $O/bench_n1_default.err
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > $O/bench_n1_long.json 2> $O/bench_n1_long.err
python benchmarks/run_configs.py --configs 1,2,4,5 --out $O/configs.json > $O/configs.log 2>&1; echo "rc=$?" >> $O/configs.log
python benchmarks/fixed_cost.py > $O/fixed_cost.json 2> $O/fixed_cost.err
python benchmarks/encoder_bench.py > $O/encoder_bench.json 2> $O/encoder_bench.err
python benchmarks/k6_probe.py > $O/k6_probe.json 2> $O/k6_probe.err
# ncu: K1 launch list of the default bench command; encoder kernels (cross-encoder 8 x 512: GEMMs + tensor-core attention)
K1CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sweep="
$K1CMD > $O/plain_k1.log 2>&1 && \
ncu --set full --clock-control none -k regex:scan_stream -s 3 -c 2 -f -o $O/r02_k1_final2 $K1CMD > $O/ncu_k1.log 2>&1
ncu -i $O/r02_k1_final2.ncu-rep --page raw --csv > $O/r02_k1_final2_raw.csv 2>/dev/null; rm -f $O/r02_k1_final2.ncu-rep
ECMD="python benchmarks/encoder_profile.py cross 8 512"
$ECMD > $O/plain_enc.log 2>&1 && \
ncu --set full --clock-control none -k regex:"gemm_wt_kernel|attention_mma_kernel" -s 24 -c 6 -f -o $O/r02_encoder2 $ECMD > $O/ncu_enc.log 2>&1
ncu -i $O/r02_encoder2.ncu-rep --page raw --csv > $O/r02_encoder_cross_b8_s512_raw.csv 2>/dev/null; rm -f $O/r02_encoder2.ncu-rep
du -sh $O; ls $O
tail -n 3 $O/smoke.log $O/pytest_gpu.log $O/configs.log $O/bench_n1_default.err
