"""CPU: the oracle (oracle/lmcma_oracle.cpp) against the golden vectors captured from the compiled
reference (tests/golden/make_golden.py), and — where /root/reference exists — against the compiled
reference itself, bit for bit."""
import numpy as np
import pytest

from conftest import REFERENCE_PRESENT, weighted_sphere


def two_gaussians(X):   # example_lmcma.cpp:18-26
    X = np.atleast_2d(X)
    return -(4 * np.exp(-((X[:, 0] - 4) ** 2 + (X[:, 1] - 4) ** 2)) + 2 * np.exp(-((X[:, 0] - 2) ** 2 + (X[:, 1] - 2) ** 2)))


def test_rng_matches_reference_stream(po, golden):
    for seed, vec in golden["rng"].items():
        assert po.rng_uniform(int(seed), 8).tolist() == vec["uniform"]
        assert po.rng_gauss(int(seed), 16).tolist() == vec["gauss"]
    # SURVEY 8c known-answer values (seed 1)
    assert po.rng_uniform(1, 1)[0] == 0.41599935685098144
    assert po.rng_gauss(1, 1)[0] == -0.83685380259280617


def test_rank_ties_match_reference_qsort(po, golden):
    g = golden["qsort_ties"]
    v, ids = po.rank(g["in"])
    assert ids.tolist() == g["ids"]
    assert v.tolist() == g["sorted"]


def test_default_lambda(po, golden):
    for n, lam in golden["default_lambda"].items():
        o = po.OracleLMCMA(int(n), x0=np.zeros(int(n)))
        assert o.ints()["lambda"] == lam
    assert golden["default_lambda"] == {"2": 6, "10": 10, "40": 15, "400": 21, "1500": 25}


def _replay(po, g, fobj, full):
    n, lam = g["n"], g["lambda"]
    o = po.OracleLMCMA(n, x0=np.array(g["x0"]), lam=lam, sigma=g["sigma0"], seed=g["seed"],
                       lo=None if g["lo"] is None else np.array(g["lo"]), hi=None if g["hi"] is None else np.array(g["hi"]))
    d = o.doubles()
    for k, v in g["consts"].items():
        assert d[k] == v, k
    assert o.ints()["mu"] == g["mu"]
    sig = [d["sigma"]]
    for gi in range(len(g["t"])):
        X = o.array("X")
        if full:
            assert np.array_equal(X, np.array(g["gens"][gi]["X"])), gi
        f = fobj(X)
        for i in range(lam):
            assert np.array_equal(o.ask(), X[i])
            o.tell(f[i])
        sig.append(o.doubles()["sigma"])
        live = o.ints()["live"]
        assert o.int_array("t")[:live].tolist() == g["t"][gi], gi
        live_slots = o.int_array("t")[:live]                       # vec is per SLOT; dead slots hold garbage in the reference
        assert np.array_equal(o.int_array("vec")[live_slots], np.array(g["vec"][gi])[live_slots]), gi
        if full:
            rec = g["gens"][gi]
            assert o.int_array("arindex").tolist() == rec["arindex"]
            for k in ("xmean", "pc", "Nj", "Lj"):
                got = o.array(k)
                want = np.array(rec[k])
                sl = slice(None)
                if k in ("Nj", "Lj"):
                    live_slots = o.int_array("t")[:live]
                    got, want = got[live_slots], want[live_slots]
                assert np.array_equal(got, want), (gi, k)
            live_slots = o.int_array("t")[:live]
            assert np.array_equal(o.array("V")[live_slots], np.array(rec["V"])[live_slots]), gi
            assert np.array_equal(o.array("P")[live_slots], np.array(rec["P"])[live_slots]), gi
            assert o.doubles()["s"] == rec["s"]
    assert sig == g["sigma"]
    assert o.doubles()["best_f"] == g["best_f"]
    assert o.ints()["counteval"] == g["counteval"]
    assert np.array_equal(o.array("xmean"), np.array(g["xmean_final"]))


def test_oracle_replays_golden_run_n10(po, golden):
    """n = 10, lambda = m = 6, 14 generations with every intermediate state (pins slot recycling, a11/a12)."""
    _replay(po, golden["run_n10"], weighted_sphere, True)
    # sigma after generations 0..7 and slot order after generation 9
    assert np.allclose(golden["run_n10"]["sigma"][1:9],
                       [1, 0.798516, 0.72313, 0.625888, 0.507627, 0.51362, 0.488512, 0.543456], rtol=2e-6)
    assert golden["run_n10"]["t"][9] == [0, 2, 4, 1, 5, 3]


def test_oracle_replays_golden_run_n40(po, golden):
    """C1's optimiser shape: n = 40, default lambda = 15, 200 generations; BestF known answer."""
    _replay(po, golden["run_n40"], weighted_sphere, False)
    assert abs(golden["run_n40"]["best_f"] - 0.0117974507803) < 1e-12
    assert golden["run_n40"]["counteval"] == 3000


def test_oracle_replays_golden_bounds_runs(po, golden):
    _replay(po, golden["run_bounds"], two_gaussians, False)
    _replay(po, golden["run_n10_bounds"], weighted_sphere, True)
    assert np.allclose(golden["run_bounds"]["xmean_final"], [4.0, 4.0], atol=1e-2)   # example_lmcma.cpp:22


def test_parameterised_m_reduces_to_reference_rule(po):
    """m = lambda passed explicitly == m defaulted (the reference's rule)."""
    a = po.OracleLMCMA(12, x0=np.full(12, 0.3), lam=8, m=0, seed=3)
    b = po.OracleLMCMA(12, x0=np.full(12, 0.3), lam=8, m=8, seed=3)
    for _ in range(20):
        X = a.array("X")
        assert np.array_equal(X, b.array("X"))
        f = weighted_sphere(X)
        a.tell_all(f); b.tell_all(f)
    assert a.doubles() == b.doubles()
    c = po.OracleLMCMA(12, x0=np.full(12, 0.3), lam=8, m=4, seed=3)
    assert c.ints()["m"] == 4 and c.doubles()["cc"] == 0.25


def test_injected_deviates_equal_stream(po):
    """Feeding the oracle the stream's own deviates through the injection port changes nothing."""
    n, lam = 10, 6
    z = po.rng_gauss(1, lam * n * 4).reshape(4, lam, n)
    a = po.OracleLMCMA(n, x0=np.full(n, 0.5), lam=lam, seed=1)
    b = po.OracleLMCMA(n, x0=np.full(n, 0.5), lam=lam, seed=99, Z0=z[0])
    for g in range(3):
        X = a.array("X")
        assert np.array_equal(X, b.array("X"))
        f = weighted_sphere(X)
        a.tell_all(f)
        b.tell_all(f, z[g + 1])


@pytest.mark.skipif(not REFERENCE_PRESENT, reason="needs /root/reference (authoring container)")
def test_oracle_bit_identical_to_compiled_reference(po):
    for n, lam, gens, seed in ((10, 6, 60, 1), (40, 0, 120, 2), (64, 24, 40, 7)):
        r = po.RefLMCMA(n, x0=np.full(n, 0.5), lam=lam, seed=seed)
        o = po.OracleLMCMA(n, x0=np.full(n, 0.5), lam=lam, seed=seed)
        L = r.ints()["lambda"]
        for e in range(gens * L):
            xr, xo = r.ask(), o.ask()
            assert np.array_equal(xr, xo)
            f = float(weighted_sphere(xr)[0])
            r.tell(f); o.tell(f)
        sr, so = r.state(), o.state()
        for k in sr:
            assert np.array_equal(np.asarray(sr[k]), np.asarray(so[k])), k
    # with box bounds and a uniform-random start (x0 = NULL, lmcma.cpp:161-163)
    r = po.RefLMCMA(5, x0=None, lam=8, lo=np.full(5, 0.1), hi=np.full(5, 0.9), seed=4)
    o = po.OracleLMCMA(5, x0=None, lam=8, lo=np.full(5, 0.1), hi=np.full(5, 0.9), seed=4)
    for e in range(200):
        xr, xo = r.ask(), o.ask()
        assert np.array_equal(xr, xo)
        f = float(weighted_sphere(xr)[0])
        r.tell(f); o.tell(f)


@pytest.mark.skipif(not REFERENCE_PRESENT, reason="needs /root/reference (authoring container)")
def test_golden_file_is_reproducible_from_the_reference(po, golden):
    assert po.rng_gauss(1, 16, which="ref").tolist() == golden["rng"]["1"]["gauss"]
    v, ids = po.rank(golden["qsort_ties"]["in"], "ref")
    assert ids.tolist() == golden["qsort_ties"]["ids"]


def test_oracle_warm_start_is_the_same_run(po):
    """orc_lmcma_set_state (used by the C4-shaped GPU parity test to start a lambda = 8192 oracle from a state with all
    m slots live): an oracle loaded with another oracle's state and fed the same deviates / fitness continues bit for
    bit like the original, through slot recycling."""
    n, lam, m = 24, 12, 7
    rng = np.random.default_rng(3)
    a = po.OracleLMCMA(n, x0=np.full(n, 0.5), lam=lam, m=m, sigma=0.4, seed=1)
    for g in range(16):
        Z = rng.standard_normal((lam, n))
        a.tell_all(weighted_sphere(a.array("X")), Z)
    b = po.OracleLMCMA(n, x0=np.zeros(n), lam=lam, m=m, sigma=9.0, seed=5)
    b.load_state(a.state(), a.array("prev_fit"), Z)
    assert np.array_equal(a.array("X"), b.array("X"))
    for g in range(12):
        Z = rng.standard_normal((lam, n))
        f = weighted_sphere(a.array("X"))
        a.tell_all(f, Z); b.tell_all(f, Z)
        sa, sb = a.state(), b.state()
        for k in ("xmean", "pc", "V", "P", "Nj", "Lj", "t", "vec"):
            assert np.array_equal(sa[k], sb[k]), (g, k)
        assert sa["sigma"] == sb["sigma"] and sa["itr"] == sb["itr"] and sa["live"] == sb["live"]
        assert np.array_equal(a.array("X"), b.array("X"))
