import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import lmcma_path_planner_b200 as L
gens = int(sys.argv[1]); start_log = int(sys.argv[2])
dist, start, goal, lo, hi, x0 = bench.build_problem()
cmap = L.CostMap(dist, "f32")
n = 2 * bench.W
dev = L.Optimizer(n, x0=x0, lam=bench.LAM, m=bench.M, lo=lo, hi=hi, sigma0=bench.SIGMA0, seed=1000, record_z=True)
dev.attach_cost(cmap, [start], [goal], bench.W, L.LONGSAFE, 1e4)
for g in range(gens):
    dev.run(1)
    if g >= start_log:
        V, P, pc = dev.get("V")[0], dev.get("P")[0], dev.get("pc")[0]
        Nj, Lj = dev.get("Nj")[0], dev.get("Lj")[0]
        X, f = dev.get("X")[0], dev.get("fit")[0]
        t = dev.get("t")[0]
        nv = (V.astype(np.float64) ** 2).sum(1)
        print("gen %d sigma %.4g |V|max %.4g |P|max %.4g |pc|max %.4g Nj[%.3g,%.3g] Lj[%.3g,%.3g] nv[%.3g,%.3g] Xnan %d fnan %d funique %d" %
              (g, dev.get("sigma")[0], np.abs(V).max(), np.abs(P).max(), np.abs(pc).max(), Nj.min(), Nj.max(), Lj.min(), Lj.max(),
               nv.min(), nv.max(), int(np.isnan(X).sum()), int(np.isnan(f).sum()), len(np.unique(f))))
        if g >= gens - 3:
            vec = dev.get("vec")[0]
            pos_of_slot = np.argsort(t)
            nan_slots = np.where(np.isnan(V).any(1))[0]
            print("   t  ", t.tolist()); print("   vec", vec[t].tolist())
            print("   nan slots", nan_slots.tolist(), "positions", [int(pos_of_slot[s]) for s in nan_slots], "nan cols per row", [int(np.isnan(V[s]).sum()) for s in nan_slots][:8])
            print("   P nan rows", np.where(np.isnan(P).any(1))[0].tolist(), "Nj by pos", np.round(Nj[t], 5).tolist()[:12])
    if g == gens - 2:
        pre = {k: dev.get(k)[0].copy() for k in ("V", "P", "pc", "Lj", "Nj", "t", "vec")}
c1, cc, cs, tgt, K, M, mueff = dev.get("consts")
post = {k: dev.get(k)[0].copy() for k in ("V", "P", "pc", "Lj", "Nj", "t", "vec")}
m = len(post["t"]); order = post["t"]
first_stale = next((i for i in range(m) if order[i] != pre["t"][i]), m - 1)
if first_stale == 1: first_stale = 0
print("first_stale", first_stale)
V = pre["V"].astype(np.float64).copy(); P = post["P"].astype(np.float64); Lj = pre["Lj"].copy()
r = c1 / (1 - c1)
for i in range(first_stale, m):
    Av = P[order[i]].copy()
    for j in range(i):
        vj = V[order[j]]
        Av = K * Av - Lj[order[j]] * (vj @ Av) * vj
    V[order[i]] = Av
    nv = Av @ Av
    Lj[order[i]] = (1.0 / (np.sqrt(1 - c1) * nv)) * (1 - 1.0 / np.sqrt(1 + r * nv))
err = np.abs(post["V"].astype(np.float64) - V).max(1) / np.maximum(np.abs(V).max(1), 1e-30)
print("rel err by position:", np.round(err[order], 6).tolist())
print("expected |v| by position:", np.round(np.sqrt((V[order] ** 2).sum(1)), 3).tolist())
print("expected Lj by position:", Lj[order].tolist()[30:])
print("device Lj by position:", post["Lj"][order].tolist()[30:])
