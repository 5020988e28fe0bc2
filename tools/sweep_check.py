"""Teacher-forced check of k_update's recompute for a chosen first_stale pattern against a numpy restatement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lmcma_path_planner_b200 as L
n, lam, m = 400, 64, 40
target = int(sys.argv[1]) if len(sys.argv) > 1 else 15
rng = np.random.default_rng(0)
dev = L.Optimizer(n, x0=np.full(n, 0.5), lam=lam, m=m, sigma0=0.5, rng="inject")
f = lambda X: (X.astype(np.float64) ** 2 * (1 + np.arange(n))).sum(1).astype(np.float32)
dev.inject_z(rng.standard_normal((lam, n)).astype(np.float32))
for g in range(44):
    X = dev.ask_all()[0]
    dev.inject_z(rng.standard_normal((lam, n)).astype(np.float32))
    dev.tell_all(f(X))
# craft stamps: strictly increasing by 2 except a gap of 1 before position `target`
t = dev.get("t")[0].copy(); itr = int(dev.get("itr")[0])
stamps = np.arange(m) * 2 + (itr - 2 * m - 5)
stamps[target:] -= 1
vec = dev.get("vec")[0].copy(); vec[t] = stamps
dev.set("vec", vec[None])
pre = {k: dev.get(k)[0].copy() for k in ("V", "P", "pc", "Lj", "Nj", "t", "vec")}
c1, cc, cs, tgt, K, M, mueff = dev.get("consts")
X = dev.ask_all()[0]
dev.inject_z(rng.standard_normal((lam, n)).astype(np.float32))
dev.tell_all(f(X))
post = {k: dev.get(k)[0].copy() for k in ("V", "P", "pc", "Lj", "Nj", "t", "vec")}
order = post["t"]
first_stale = next(i for i in range(m) if order[i] != pre["t"][i])
print("first_stale observed", first_stale, "(1 -> 0 rule applies)" )
if first_stale == 1: first_stale = 0
# numpy recompute in FP64 (lmcma.cpp:373-390, 449-463)
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
print("rel err by position:", np.round(err[order], 7).tolist())
print("nan rows:", np.where(np.isnan(post["V"]).any(1))[0].tolist())
