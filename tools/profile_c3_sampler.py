"""A small C3-shaped batch (queries x lambda 64, n = 400, m = 40) for `ncu -k regex:k_sample`: 45 fused generations to
fill the pairs, then two un-graphed generations (profile_kernels) whose sampler launches are the ones to capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lmcma_path_planner_b200 as L
from lmcma_path_planner_b200 import maps
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
W, M = 200, 40
dist, _, _ = maps.config2_map(size=1024, n_rects=128, seed=42, clamp=64.0)
starts, goals = maps.random_queries(dist, Q, seed=7, min_sep=256)
lo, hi = maps.box_bounds((1024, 1024), W)
x0 = np.stack([maps.straight_line(starts[q], goals[q], W) for q in range(Q)])
cm = L.CostMap(dist, "f32")
opt = L.Optimizer(2 * W, x0=x0, lam=64, m=M, batch=Q, lo=lo, hi=hi, sigma0=8.0, seed=7)
opt.attach_cost(cm, starts, goals, W, L.LONGSAFE, 1e4)
opt.run(45)
opt.sync()
print(opt.profile_kernels(2))
