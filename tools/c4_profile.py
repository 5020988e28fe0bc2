"""ncu driver: C4-shaped instance (3-D u8 map, n = 1500), a few un-graphed generations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lmcma_path_planner_b200 as L
from lmcma_path_planner_b200 import maps
size, W, lam = 128, 500, 1024
n = 3 * W; m = int(2 * np.sqrt(n))
dmap, start, goal = maps.config4_map(size=size, n_boxes=max(64, 4096 * size ** 3 // 512 ** 3), seed=43)
cmap = L.CostMap(dmap, "u8", u8_scale=0.25)
lo, hi = maps.box_bounds((size, size, size), W)
opt = L.Optimizer(n, x0=maps.straight_line(start, goal, W), lam=lam, m=m, lo=lo, hi=hi, sigma0=8.0 * size / 512, seed=43)
opt.attach_cost(cmap, [start], [goal], W, L.LONGSAFE, 1e4)
for g in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12): opt.profile_kernels(1)
print(opt.profile_kernels(3))
