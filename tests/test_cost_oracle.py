"""CPU: the declared cost model (oracle/cost_oracle.c) — analytic known answers, the distance-field
restatements, and the bundled maps."""
import os

import numpy as np
import pytest

from conftest import REFERENCE_PRESENT
from lmcma_path_planner_b200 import maps


def test_straight_line_on_empty_map_known_answer(po):
    dist = np.full((100, 100), 10.0, np.float32)            # clearance 10 everywhere
    W, start, goal = 9, (10.0, 10.0), (10.0, 60.0)
    x = maps.straight_line(start, goal, W)
    r = po.CostProblem(dist, start, goal, W, 1.0, 1.0, 1e4).evaluate(x)
    assert r["ncoll"][0] == 0
    assert abs(r["length"][0] - 50.0) < 1e-9
    assert abs(r["clearance"][0] - 50.0 / 10.0) < 1e-9       # integral of 1/10 over length 50
    assert r["nsamp"][0] == 10 * (5 + 1)                      # 10 segments of 5 cells -> K = 5
    assert abs(r["f"][0] - 55.0) < 1e-9


def test_collisions_counted_once_per_sample(po):
    dist = np.full((50, 50), 5.0, np.float32)
    dist[:, 20:23] = 0.0                                      # a 3-cell wall across every row
    W, start, goal = 4, (0.0, 10.0), (40.0, 10.0)
    x = maps.straight_line(start, goal, W)                    # horizontal, 5 segments of 8 cells
    r = po.CostProblem(dist, start, goal, W, 1.0, 0.0, 1.0, c_min=0.5).evaluate(x)
    assert r["ncoll"][0] == 3                                 # cells x = 20, 21, 22 once each
    assert abs(r["f"][0] - (40.0 + 3.0)) < 1e-9
    # outside the map is a collision (declared)
    x2 = x.copy(); x2[1] = -5.0
    r2 = po.CostProblem(dist, start, goal, W).evaluate(x2)
    assert r2["ncoll"][0] > 3


def test_dimension_major_layout(po):
    """x[d*W + w] (lmcma.cpp:786-791): swapping the two halves swaps x and y."""
    dist = np.full((64, 64), 3.0, np.float32)
    dist[40:, :] = 0.0                                        # rows (y) >= 40 blocked
    W = 3
    xs, ys = np.array([10.0, 20.0, 30.0]), np.array([5.0, 5.0, 5.0])
    a = po.CostProblem(dist, (0, 5), (40, 5), W).evaluate(np.concatenate([xs, ys]))
    assert a["ncoll"][0] == 0
    b = po.CostProblem(dist, (5, 0), (5, 63), W).evaluate(np.concatenate([ys, np.array([20.0, 45.0, 60.0])]))
    assert b["ncoll"][0] > 0


def test_rint_half_to_even_cells(po):
    dist = np.arange(100, dtype=np.float32).reshape(10, 10) + 1
    W = 1
    prob = po.CostProblem(dist, (0.5, 0.0), (2.5, 0.0), W)     # samples at x = 0.5, 1.5, 2.5 -> cells 0, 2, 2
    cells = prob.trace(np.array([1.5, 0.0], np.float32))
    assert cells.tolist() == [0, 2, 2, 2]


def test_exact_edt_matches_scipy(po):
    from scipy import ndimage
    rng = np.random.default_rng(0)
    occ = (rng.random((37, 53)) < 0.05).astype(np.uint8)
    assert np.allclose(po.edt_exact(occ), ndimage.distance_transform_edt(occ == 0), atol=1e-5)
    occ3 = (rng.random((9, 11, 13)) < 0.03).astype(np.uint8)
    assert np.allclose(po.edt_exact(occ3), ndimage.distance_transform_edt(occ3 == 0), atol=1e-5)
    assert np.allclose(po.edt_exact(occ, clamp=3.0), np.minimum(ndimage.distance_transform_edt(occ == 0), 3.0), atol=1e-5)
    assert np.allclose(maps.distance_field(occ), po.edt_exact(occ), atol=1e-5)


def test_8ssedt_restatement(po, golden_maps):
    """The restated 8SSEDT (planner.cpp:403-490): exact on obstacle cells, never below the true distance, within
    the method's known small over-estimate elsewhere; the signed variant follows planner.cpp:536-540."""
    from scipy import ndimage
    occ = golden_maps["problem1"]
    sq = po.ssedt8_sq(occ)
    true = ndimage.distance_transform_edt(occ == 0)
    d = np.sqrt(sq.astype(np.float64))
    assert np.all(d[occ == 1] == 0)
    assert np.all(d >= true - 1e-9)
    assert np.max(d - true) < 1.0
    sd = po.ssedt8_signed(occ)
    assert np.all(sd[occ == 1] <= 0) and np.all(sd[occ == 0] >= 0)
    inner = ndimage.distance_transform_edt(occ == 1)
    assert np.array_equal(sd[occ == 0], np.sqrt(sq[occ == 0].astype(np.float64)).astype(int))
    assert np.all(np.abs(-sd[occ == 1] - inner[occ == 1].astype(int)) <= 1)


def test_bundled_maps_match_their_description(golden_maps):
    assert np.array_equal(golden_maps["problem1"], maps.problem_occupancy(1))
    assert np.array_equal(golden_maps["problem2"], maps.problem_occupancy(2))
    assert np.array_equal(golden_maps["two_bars"], maps.two_bars_occupancy())


@pytest.mark.skipif(not REFERENCE_PRESENT, reason="needs /root/reference (authoring container)")
def test_golden_maps_match_reference_bmps(golden_maps):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for name in ("problem1", "problem2"):
        img = mg.read_bmp24(os.path.join(mg.REF_IMAGES, name + ".bmp"))
        assert set(np.unique(img)) <= {0, 255}
        assert np.array_equal((img[:, :, 1] < 128).astype(np.uint8), golden_maps[name])


def test_threads_do_not_change_results(po):
    dist, start, goal = maps.config2_map(size=256, n_rects=16, seed=1, clamp=32.0)
    W = 30
    rng = np.random.default_rng(1)
    X = (maps.straight_line(start, goal, W)[None] + 5 * rng.standard_normal((37, 2 * W))).astype(np.float32)
    a = po.CostProblem(dist, start, goal, W, threads=1).evaluate(X)
    b = po.CostProblem(dist, start, goal, W, threads=5).evaluate(X)
    for k in a:
        assert np.array_equal(a[k], b[k])
