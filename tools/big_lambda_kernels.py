"""Per-kernel times of one generation at the C5 population sizes (lambda = 1024 * 2^r, n = 400, m = 40, cluttered C2 map):
is the O(lambda^2) rank-by-counting still small next to the cost kernel at lambda = 65536?
  python tools/big_lambda_kernels.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lmcma_path_planner_b200 as L
from lmcma_path_planner_b200 import maps

W, M = 200, 40
dist, start, goal = maps.config2_map(n_rects=4 * 2048, seed=42)
lo, hi = maps.box_bounds((4096, 4096), W)
cm = L.CostMap(dist, "f32")
x0 = maps.straight_line(start, goal, W)
for lam in (1024, 4096, 8192, 16384, 32768, 65536):
    opt = L.Optimizer(2 * W, x0=x0, lam=lam, m=M, lo=lo, hi=hi, sigma0=32.0, seed=1)
    opt.attach_cost(cm, [start], [goal], W, L.LONGSAFE, 1e4)
    opt.run(45)
    opt.sync()
    ms = opt.last_run_ms() / 45
    for _ in range(2):
        pk = opt.profile_kernels(2)
    tot = sum(pk.values())
    print("lambda %6d: fused generation %.4f ms | kernels (serial, events) %s | rank share %.1f %%" %
          (lam, ms, {k: round(v, 4) for k, v in pk.items()}, 100 * pk["rank"] / tot), flush=True)
    opt.close()
