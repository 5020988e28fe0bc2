"""Free-running device LM-CMA vs the FP64 oracle fed the device's own deviates and fitness, on the C2 workload.
Prints sigma of both every few generations: they should track for tens of generations and stay qualitatively alike."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import lmcma_path_planner_b200 as L
from oracle import pyoracle as po

po.build()
gens = int(sys.argv[1]) if len(sys.argv) > 1 else 120
lam = int(sys.argv[2]) if len(sys.argv) > 2 else bench.LAM
dist, start, goal, lo, hi, x0 = bench.build_problem()
cmap = L.CostMap(dist, "f32")
n = 2 * bench.W
dev = L.Optimizer(n, x0=x0, lam=lam, m=bench.M, lo=lo, hi=hi, sigma0=bench.SIGMA0, seed=1000, record_z=True)
dev.attach_cost(cmap, [start], [goal], bench.W, L.LONGSAFE, 1e4)
ora = po.OracleLMCMA(n, x0=x0, lam=lam, m=bench.M, lo=lo, hi=hi, sigma=bench.SIGMA0, seed=1, Z0=dev.get("Z")[0].astype(np.float64))
for g in range(gens):
    Xd = dev.get("X")[0]
    Xo = ora.array("X")
    dev.run(1)                                   # cost -> rank -> update -> sample on the device
    f = dev.get("fit")[0]
    ora.tell_all(f.astype(np.float64), dev.get("Z")[0].astype(np.float64))
    if g % 10 == 0 or g == gens - 1:
        sd, so = float(dev.get("sigma")[0]), ora.doubles()["sigma"]
        print("gen %3d sigma dev %.6g ora %.6g | max|X dev - X ora| %.3g | mean f %.4g nsamp %.0f" %
              (g, sd, so, float(np.abs(Xd - Xo).max()), float(f.mean()), float(dev.get("nsamp").mean())))
