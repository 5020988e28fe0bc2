"""Timing of the smoothness-prior sampling path (k_gauss + k_prior: Z L^T for the whole population, the one dense contraction on
the path) at the C2 shape, next to the same sampler without the prior.  python tools/prior_timing.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lmcma_path_planner_b200 as L
from lmcma_path_planner_b200 import maps, _capi as K

W, lam, m = 200, 1024, 40
n = 2 * W
dist, start, goal = maps.config2_map(size=1024, n_rects=128, seed=42, clamp=64.0)
lo, hi = maps.box_bounds((1024, 1024), W)
cm = L.CostMap(dist, "f32")
cov = np.zeros(n * n)
assert K.lib().lmcma_b200_covariance(2, W, K.dptr(cov)) == 0
for label, c in (("no prior", None), ("covariance(2, 200) prior", cov.reshape(n, n))):
    opt = L.Optimizer(n, x0=maps.straight_line(start, goal, W), lam=lam, m=m, lo=lo, hi=hi, sigma0=8.0, seed=1, covariance=c)
    opt.attach_cost(cm, [start], [goal], W, L.LONGSAFE, 1e4)
    opt.run(45)
    for _ in range(2):
        pk = opt.profile_kernels(5)
    flop = 2.0 * lam * n * n / 2            # lower-triangular L: half of lambda x n x n multiply-adds
    print("%-26s sample stage %.4f ms (k_gauss + k_prior + sampler)  %s" % (label, pk["sample"], {k: round(v, 4) for k, v in pk.items()}))
    if c is not None:
        print("  prior contraction: %.3g FLOP -> the extra %.1f us" % (flop, 1e3 * (pk["sample"] - base)))
    else:
        base = pk["sample"]
    opt.close()
