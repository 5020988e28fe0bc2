"""the reference's sibling optimisers (SepCMA lmcma.hpp:152, CMAChol lmcma.hpp:211) as CPU
cross-checks of solution quality.  tests/golden/siblings_reference.json holds what the UNMODIFIED reference
(oracle/_ref) reached on fixed problems (tests/golden/make_siblings_golden.py); here the LM-CMA restatement
(oracle/lmcma_oracle.cpp, the checker of every GPU parity test) replays the LMCMA entry bit for bit and is compared
with the siblings.  No GPU, no /root/reference needed for the replay."""
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import REFERENCE_PRESENT, ROOT, weighted_sphere

GOLDEN = os.path.join(ROOT, "tests", "golden", "siblings_reference.json")


@pytest.fixture(scope="module")
def siblings():
    with open(GOLDEN) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def maker():
    spec = importlib.util.spec_from_file_location("make_siblings_golden",
                                                  os.path.join(ROOT, "tests", "golden", "make_siblings_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def oracle_best(o, generations, cost):
    """ask / tell through the restatement, BestF rule of lmcma.cpp:190-193 (strict <, first evaluation included)."""
    lam = o.ints()["lambda"]
    best, bx = None, None
    for _ in range(generations):
        if o.done():
            break
        for _ in range(lam):
            x = o.ask()
            f = cost(x)
            if best is None or f < best:
                best, bx = f, x.copy()
            o.tell(f)
    return best, bx


def test_oracle_replays_the_reference_lmcma_entries(po, siblings, maker):
    for c in siblings["sphere"]:
        o = po.OracleLMCMA(c["n"], x0=np.full(c["n"], 0.5), lam=c["lam"], sigma=c["sigma"], seed=c["seed"])
        assert o.ints()["lambda"] == c["lambda"]
        best, _ = oracle_best(o, c["generations"], lambda x: float(weighted_sphere(x)[0]))
        assert best == c["best"]["LMCMA"]
    c = siblings["plan"]
    prob, x0, lo, hi = maker.plan_problem(c)
    assert float(prob.evaluate(x0)["f"][0]) == c["start_cost"]
    o = po.OracleLMCMA(prob.n, x0=x0, lam=c["lam"], lo=lo, hi=hi, sigma=c["sigma"], seed=c["seed"])
    best, bx = oracle_best(o, c["generations"], lambda x: float(prob.evaluate(x)["f"][0]))
    assert best == c["best"]["LMCMA"]
    assert int(prob.evaluate(bx)["ncoll"][0]) == c["ncoll"]["LMCMA"] == 0


def test_lmcma_solution_quality_against_the_siblings(siblings):
    """The three optimisers agree on where the optimum is: on the convex case all are within the same budget far
    below the start (f(x0) = 0.25 * n(n+1)/2), on the planning case all three paths are collision-free and LM-CMA's
    cost is within 5 % of the better sibling (start cost 8.9e4 with collisions)."""
    for c in siblings["sphere"]:
        f0 = 0.25 * c["n"] * (c["n"] + 1) / 2
        for k, v in c["best"].items():
            assert 0.0 <= v < 1e-4 * f0, (k, v)
    p = siblings["plan"]
    assert all(v == 0 for v in p["ncoll"].values())
    assert all(v < 0.1 * p["start_cost"] for v in p["best"].values())
    assert p["best"]["LMCMA"] <= 1.05 * min(p["best"]["SepCMA"], p["best"]["CMAChol"])


@pytest.mark.skipif(not REFERENCE_PRESENT, reason="needs /root/reference (authoring container)")
def test_siblings_golden_is_reproducible_from_the_reference(po, siblings, maker):
    assert json.loads(json.dumps(maker.run_all())) == siblings
