"""Writes tests/golden/siblings_reference.json: solution quality of the reference's three optimisers (LMCMA and its
siblings SepCMA / CMAChol, lmcma.hpp:131-254) on fixed problems, run HERE from the unmodified reference compiled into
oracle/_ref (needs /root/reference).  SURVEY 8f.4: the siblings are CPU cross-checks, never GPU targets.

    python tests/golden/make_siblings_golden.py

Problems: the weighted sphere of tests/conftest.py (n = 10, 40) and the C1-shaped planning problem (two_bars map,
100 x 100, W = 20 waypoints, box bounds, straight-line start) under the declared cost model (oracle/cost_oracle.c).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import weighted_sphere  # noqa: E402
from lmcma_path_planner_b200 import maps  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

KINDS = ("LMCMA", "SepCMA", "CMAChol")
SPHERE_CASES = ({"n": 10, "lam": 0, "generations": 150, "seed": 3, "sigma": 0.3},
                {"n": 40, "lam": 0, "generations": 300, "seed": 5, "sigma": 0.3})
PLAN_CASE = {"map": "two_bars", "start": [5.0, 5.0], "goal": [94.0, 94.0], "waypoints": 20, "lam": 0,
             "generations": 400, "seed": 11, "sigma": 5.0}


def plan_problem(case):
    occ = dict(np.load(os.path.join(ROOT, "tests", "golden", "maps_2d.npz")))[case["map"]]
    dist = maps.distance_field(occ)
    prob = po.CostProblem(dist, case["start"], case["goal"], case["waypoints"])
    x0 = maps.straight_line(case["start"], case["goal"], case["waypoints"]).astype(np.float64)
    lo, hi = maps.box_bounds(dist.shape[::-1], case["waypoints"])
    return prob, x0, np.asarray(lo, np.float64), np.asarray(hi, np.float64)


def run_all():
    out = {"sphere": [], "plan": None}
    for c in SPHERE_CASES:
        row = dict(c, best={})
        for k in KINDS:
            o = po.RefSibling(k, c["n"], x0=np.full(c["n"], 0.5), lam=c["lam"], sigma=c["sigma"], seed=c["seed"])
            row["lambda"] = o.lam
            row["best"][k] = o.run(c["generations"], func=weighted_sphere)[0]
        out["sphere"].append(row)
    c = PLAN_CASE
    prob, x0, lo, hi = plan_problem(c)
    row = dict(c, best={}, ncoll={}, start_cost=float(prob.evaluate(x0)["f"][0]))
    for k in KINDS:
        o = po.RefSibling(k, prob.n, x0=x0, lam=c["lam"], lo=lo, hi=hi, sigma=c["sigma"], seed=c["seed"])
        row["lambda"] = o.lam
        best, bx = o.run(c["generations"], problem=prob)
        row["best"][k] = best
        row["ncoll"][k] = int(prob.evaluate(bx)["ncoll"][0])
    out["plan"] = row
    return out


if __name__ == "__main__":
    po.build()
    res = run_all()
    with open(os.path.join(ROOT, "tests", "golden", "siblings_reference.json"), "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))
