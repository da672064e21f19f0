#!/bin/bash
set -u
O=gpurun_out/r2st3
mkdir -p $O
P=multimodal-rag-for-image-text-search_b200
for rep in 1 2; do
for v in old new; do
  cp benchmarks/_ab/store_$v.py $P/store.py; cp benchmarks/_ab/hosttable_$v.py $P/hosttable.py
  python benchmarks/store_overhead.py > $O/ov_${v}_$rep.json 2>> $O/err.log
  echo "$v $rep $(cat $O/ov_${v}_$rep.json)"
done
done
tail -2 $O/err.log
