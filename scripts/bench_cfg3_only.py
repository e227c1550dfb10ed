"""BASELINE config 3 alone (bench.secondary_cfg3 on one GPU): A/B runs of knobs that only move the Kronecker CG."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cola_b200 as cb
torch.cuda.set_device(0)
ctx = bench.DistCtx()
r = bench.secondary_cfg3(ctx, cb)
print(json.dumps({"iters_per_s": r["iters_per_s"], "ms_per_iter": r["ms_per_iter"], "frac": r["roofline"]["frac"], "matmat_ms": r["roofline"]["matmat"]["ms"]}))
