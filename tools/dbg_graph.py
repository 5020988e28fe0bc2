import os, sys
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo") if os.path.isdir("/root/repo") else None
import numpy as np, bench
import lmcma_path_planner_b200 as L
dist, start, goal, lo, hi, x0 = bench.build_problem()
cmap = L.CostMap(dist, "f32")
opt = L.Optimizer(2 * bench.W, x0=x0, lam=bench.LAM, m=bench.M, lo=lo, hi=hi, sigma0=bench.SIGMA0, seed=1000)
opt.attach_cost(cmap, [start], [goal], bench.W, L.LONGSAFE, 1e4)
opt.run(60); opt.sync()
for _ in range(3):
    opt.run(1); opt.sync()
