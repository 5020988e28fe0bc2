"""Debug: 3-D cost parity flakiness at 32-thread CTAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lmcma_path_planner_b200 as L
from lmcma_path_planner_b200 import maps
from oracle import pyoracle as po
po.build()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dist, start, goal = maps.config4_map(size=64, n_boxes=4096, seed=9, clamp=16.0)
W = 40
rng = np.random.default_rng(4)
lo, hi = maps.box_bounds((64, 64, 64), W)
x0 = maps.straight_line(start, goal, W)
X = np.clip(x0[None] + 3.0 * rng.standard_normal((200, 3 * W)), lo, hi).astype(np.float32)
Xz = np.zeros_like(X)
refs = {}
bad = 0
for rep in range(reps):
    for tpt in (0, 32, 64):
        if tpt: os.environ["LMCMA_B200_COST_TPT"] = str(tpt)
        else: os.environ.pop("LMCMA_B200_COST_TPT", None)
        for storage in ("f32", "u8"):
            cm = L.CostMap(dist, storage, u8_scale=0.125)
            dd = dist if storage == "f32" else cm.dequantized()
            prob = po.CostProblem(dd, start, goal, W)
            if storage not in refs:
                refs[storage] = prob.evaluate(X)
                refs[storage + "z"] = prob.evaluate(Xz)
            ref = refs[storage]
            got = cm.evaluate(X, start, goal, W)
            ok = np.array_equal(got["ncoll"], ref["ncoll"]) and np.array_equal(got["nsamp"], ref["nsamp"])
            if not ok:
                bad += 1
                rz = refs[storage + "z"]
                print("BAD rep", rep, tpt, storage, "ncoll", got["ncoll"][:5], "nsamp", got["nsamp"][:5], "ref nsamp", ref["nsamp"][:5], "f", got["f"][:3], ref["f"][:3],
                      "zero-X ref: ncoll", rz["ncoll"][:2], "nsamp", rz["nsamp"][:2], "f", rz["f"][:2], flush=True)
                got2 = cm.evaluate(X, start, goal, W)
                print("   second call ok:", np.array_equal(got2["ncoll"], ref["ncoll"]))
            tr_g, tr_r = cm.trace(X[0], start, goal, W), prob.trace(X[0])
            if not np.array_equal(tr_g, tr_r): print("BAD trace", rep, tpt, storage)
print("bad", bad)
