"""Profiling driver (run under ncu on the GPU box): the C2 workload advanced past generation m so that all
40 direction pairs are live, then a few more generations.  Prints CUDA-event per-kernel timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import lmcma_path_planner_b200 as L  # noqa: E402

storage = sys.argv[1] if len(sys.argv) > 1 else "f32"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 45
dist, start, goal, lo, hi, x0 = bench.build_problem()
cmap = L.CostMap(dist, storage)
opt = L.Optimizer(2 * bench.W, x0=x0, lam=bench.LAM, m=bench.M, lo=lo, hi=hi, sigma0=bench.SIGMA0, seed=1000)
opt.attach_cost(cmap, [start], [goal], bench.W, L.LONGSAFE, 1e4)
for g in range(warm):
    opt.profile_kernels(1)          # un-graphed launches
pk = opt.profile_kernels(5)
print("per-kernel ms (L2 warm, events):", {k: round(v, 5) for k, v in pk.items()})
print("mean samples/trajectory:", float(opt.get("nsamp").mean()), "sigma:", float(opt.get("sigma")[0]),
      "mean collisions:", float(opt.get("ncoll").mean()), "best f:", float(opt.best()[1][0]))
