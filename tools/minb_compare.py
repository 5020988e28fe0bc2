"""k_cost on the C2 shape with LMCMA_B200_COST_MINB = 7 (default, 32 registers) against 6 (40 registers, no spill):
CUDA events around the un-graphed kernel (lmcma_b200_profile_kernels), alternating, L2 warm (no flush: torch-free so
that the whole script fits a few seconds of GPU time).  DESIGN.md section 8.7."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmcma_path_planner_b200 as L  # noqa: E402
from lmcma_path_planner_b200 import maps  # noqa: E402

W, LAM, M, SIGMA0, SIZE = 200, 1024, 40, 32.0, 4096
t0 = time.time()
start, goal = (64.0, 64.0), (SIZE - 64.0, SIZE - 64.0)
occ = maps.random_boxes_occupancy((SIZE, SIZE), 2048, 8, 128, 42, border=2, clear=[(start, 64.0), (goal, 64.0)])
dist = maps.edt_device(occ, 256.0)
lo, hi = maps.box_bounds((SIZE, SIZE), W)
cm = L.CostMap(dist, "f32")
opt = L.Optimizer(2 * W, x0=maps.straight_line(start, goal, W), lam=LAM, m=M, lo=lo, hi=hi, sigma0=SIGMA0, seed=1000, rng="philox")
opt.attach_cost(cm, [start], [goal], W, L.LONGSAFE, 1e4)
opt.run(20)
print("setup %.1f s" % (time.time() - t0), flush=True)
res = {"7": [], "6": []}
for rnd in range(6):
    for minb in ("7", "6"):
        os.environ["LMCMA_B200_COST_MINB"] = minb
        opt.profile_kernels(2)
        res[minb].append(np.mean([opt.profile_kernels(1)["cost"] for _ in range(20)]))
for k, v in res.items():
    print("MINB=%s k_cost ms per launch (6 rounds of 20): %s  median %.4f" % (k, " ".join("%.4f" % x for x in v), float(np.median(v))))
print("mean samples per trajectory %.1f" % float(opt.get("nsamp").mean()))
