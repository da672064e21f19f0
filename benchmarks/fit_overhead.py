"""Scan time vs table size at batch 1 (K1) and 8 (K2): the intercept of the linear fit is the fixed cost per search."""
import importlib, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg=importlib.import_module("multimodal-rag-for-image-text-search_b200")
dev=torch.device('cuda:0')
ix=bench.build_shard(pkg,0,10_000_000,512,'bf16',dev)
q=torch.from_numpy(bench.gen_queries(1,512)).cuda()
q8=torch.from_numpy(bench.gen_queries(8,512)).cuda()
for n in (156250, 312500,625000,1250000,2500000,5000000,10000000):
    sub=pkg.ResidentIndex(ix.rows[:n])
    for name,qq in (('B1',q),('B8',q8)):
        for _ in range(5): sub.search(qq,10)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        reps=200
        e0.record()
        for _ in range(reps): sub.search(qq,10)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/reps
        print(n,name,round(ms*1000,1),'us', round(n*1024/ms/1e6,1),'GB/s')
