"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU checkers in oracle/.

* ``OracleLMCMA``  — oracle/lmcma_oracle.cpp, our FP64 restatement (parameterised in m).
* ``RefLMCMA``     — oracle/_ref/libref_lmcma.so, the UNMODIFIED reference optimiser
  (/root/reference/lmcma_path_planner/src/lmcma.cpp) behind oracle/ref_harness.cpp.
* ``CostProblem``  — oracle/cost_oracle.c, the declared trajectory-cost model.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_c_double_p = C.POINTER(C.c_double)
_c_float_p = C.POINTER(C.c_float)
_c_int_p = C.POINTER(C.c_int)
_c_i64_p = C.POINTER(C.c_int64)


def build(force=False):
    """Compile liboracle.so (always possible) and _ref/libref_lmcma.so (needs /root/reference)."""
    lib = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(lib) or not os.path.exists(os.path.join(HERE, "_ref", "libref_lmcma.so")) or \
            not os.path.exists(os.path.join(HERE, "_ref", "libref_cost.so")):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    return lib


def _dp(a):
    return a.ctypes.data_as(_c_double_p) if a is not None else None


def _fp(a):
    return a.ctypes.data_as(_c_float_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_c_int_p) if a is not None else None


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.orc_lmcma_create.restype = C.c_void_p
        L.orc_lmcma_create.argtypes = [C.c_int, C.c_int, C.c_int, _c_double_p, _c_double_p, _c_double_p,
                                       C.c_double, C.c_long, _c_double_p]
        L.orc_lmcma_destroy.argtypes = [C.c_void_p]
        L.orc_lmcma_ask.argtypes = [C.c_void_p, _c_double_p]
        L.orc_lmcma_tell.argtypes = [C.c_void_p, C.c_double]
        L.orc_lmcma_tell_all.argtypes = [C.c_void_p, _c_double_p, _c_double_p]
        L.orc_lmcma_done.argtypes = [C.c_void_p]
        L.orc_lmcma_set_state.argtypes = [C.c_void_p] + [_c_double_p] * 7 + [_c_int_p, _c_int_p, C.c_int, C.c_int,
                                                                              C.c_double, C.c_double, _c_double_p]
        L.orc_lmcma_get_ints.argtypes = [C.c_void_p, _c_int_p]
        L.orc_lmcma_get_doubles.argtypes = [C.c_void_p, _c_double_p]
        L.orc_lmcma_get_array.argtypes = [C.c_void_p, C.c_int, _c_double_p]
        L.orc_lmcma_get_int_array.argtypes = [C.c_void_p, C.c_int, _c_int_p]
        L.orc_rng_uniform.argtypes = [C.c_long, C.c_int, _c_double_p]
        L.orc_rng_gauss.argtypes = [C.c_long, C.c_long, C.c_long, _c_double_p]
        L.orc_rank.argtypes = [C.c_int, _c_double_p, _c_int_p]
        L.orc_cost_batch.argtypes = [C.c_void_p, _c_float_p, C.c_int, _c_double_p, _c_int_p, _c_int_p,
                                     _c_double_p, _c_double_p]
        L.orc_cost_sample.argtypes = [C.c_void_p, _c_float_p, C.c_int, _c_i64_p, _c_int_p, _c_double_p]
        L.orc_cost_trace.restype = C.c_int64
        L.orc_cost_trace.argtypes = [C.c_void_p, _c_float_p, _c_i64_p, C.c_int64]
        L.orc_edt_exact.argtypes = [C.c_void_p, C.c_int, _c_int_p, C.c_float, _c_float_p]
        L.orc_8ssedt_sq.argtypes = [C.c_void_p, C.c_int, C.c_int, _c_int_p]
        L.orc_8ssedt_signed.argtypes = [C.c_void_p, C.c_int, C.c_int, _c_int_p]
        _lib = L
    return _lib


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libref_lmcma.so")) or \
        os.path.exists("/root/reference/lmcma_path_planner/src/lmcma.cpp")


_ref_o0 = None


def ref(o0=False):
    """The compiled reference: -O2 (default) or the -O0 -g build (the reference's CMakeLists set CMAKE_BUILD_TYPE Debug)."""
    global _ref, _ref_o0
    if o0:
        if _ref_o0 is None:
            build()
            _ref_o0 = _bind_ref(C.CDLL(os.path.join(HERE, "_ref", "libref_lmcma_O0.so")))
        return _ref_o0
    if _ref is None:
        build()
        _ref = _bind_ref(C.CDLL(os.path.join(HERE, "_ref", "libref_lmcma.so")))
    return _ref


def _bind_ref(R):
    R.ref_lmcma_create.restype = C.c_void_p
    R.ref_lmcma_create.argtypes = [_c_double_p, C.c_int, C.c_int, _c_double_p, _c_double_p, C.c_double,
                                   _c_double_p, C.c_int]
    R.ref_lmcma_destroy.argtypes = [C.c_void_p]
    R.ref_lmcma_ask.argtypes = [C.c_void_p, _c_double_p]
    R.ref_lmcma_tell.argtypes = [C.c_void_p, C.c_double]
    R.ref_lmcma_done.argtypes = [C.c_void_p]
    R.ref_lmcma_get_ints.argtypes = [C.c_void_p, _c_int_p]
    R.ref_lmcma_get_doubles.argtypes = [C.c_void_p, _c_double_p]
    R.ref_lmcma_get_array.argtypes = [C.c_void_p, C.c_int, _c_double_p]
    R.ref_lmcma_get_int_array.argtypes = [C.c_void_p, C.c_int, _c_int_p]
    R.ref_lmcma_generation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _c_double_p, _c_double_p]
    if hasattr(R, "ref_sibling_create"):          # a prebuilt _ref from before the siblings were exposed lacks them
        R.ref_sibling_create.restype = C.c_void_p
        R.ref_sibling_create.argtypes = [C.c_int, _c_double_p, C.c_int, C.c_int, _c_double_p, _c_double_p,
                                         C.c_double, C.c_int]
        R.ref_sibling_destroy.argtypes = [C.c_void_p]
        R.ref_sibling_lambda.argtypes = [C.c_void_p]
        R.ref_sibling_run.restype = C.c_double
        R.ref_sibling_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _c_double_p]
    R.ref_rng_uniform.argtypes = [C.c_long, C.c_int, _c_double_p]
    R.ref_rng_gauss.argtypes = [C.c_long, C.c_long, C.c_long, _c_double_p]
    R.ref_myqsort.argtypes = [C.c_int, _c_double_p, _c_int_p]
    R.ref_covariance.argtypes = [C.c_int, C.c_int, _c_double_p]
    return R


_refcost = None


def ref_cost_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libref_cost.so"))


class RefCostPieces:
    """The reference's own ValidityChecker / ClearanceObjective / shortrisky / longsafe (planner.cpp:587-690), compiled
    against the OMPL stand-in by `make -C oracle ref_cost`.  One global map (the reference keeps EDT_Matrix global)."""

    def __init__(self, dist):
        global _refcost
        if _refcost is None:
            build()
            R = C.CDLL(os.path.join(HERE, "_ref", "libref_cost.so"))
            R.ref_cost_set_map.argtypes = [_c_double_p, C.c_int, C.c_int]
            R.ref_cost_is_valid.argtypes = [C.c_double, C.c_double]
            R.ref_cost_clearance.restype = C.c_double
            R.ref_cost_clearance.argtypes = [C.c_double, C.c_double]
            R.ref_cost_state_cost.restype = C.c_double
            R.ref_cost_state_cost.argtypes = [C.c_double, C.c_double]
            R.ref_cost_weights.argtypes = [C.c_int, _c_double_p]
            R.ref_cost_threshold_path_length.restype = C.c_double
            _refcost = R
        self._R = _refcost
        self.set_map(dist)

    def set_map(self, dist):
        d = np.ascontiguousarray(dist, np.float64)
        self.shape = d.shape
        self._R.ref_cost_set_map(_dp(d), d.shape[0], d.shape[1])

    def is_valid(self, x, y):
        return bool(self._R.ref_cost_is_valid(float(x), float(y)))

    def clearance(self, x, y):
        return float(self._R.ref_cost_clearance(float(x), float(y)))

    def state_cost(self, x, y):
        return float(self._R.ref_cost_state_cost(float(x), float(y)))

    def weights(self, kind):
        out = np.zeros(2, np.float64)
        interp = self._R.ref_cost_weights({"shortrisky": 0, "longsafe": 1}[kind], _dp(out))
        return float(out[0]), float(out[1]), interp

    def threshold_path_length(self):
        return float(self._R.ref_cost_threshold_path_length())


ARRAYS = {"xmean": 0, "xold": 1, "pc": 2, "V": 3, "P": 4, "Nj": 5, "Lj": 6, "X": 7, "fit": 8,
          "prev_fit": 9, "weights": 10}
INT_ARRAYS = {"t": 0, "vec": 1, "iterator": 2, "arindex": 3}
INTS = ["n", "lambda", "mu", "itr", "sample_idx", "counteval", "m", "maxsteps", "live"]
DOUBLES = ["sigma", "s", "c1", "cc", "cs", "target", "K", "M", "mueff", "best_f"]


class _Base:
    _prefix = None

    def _fn(self, name):
        return getattr(self._L, self._prefix + name)

    def ints(self):
        out = np.zeros(9, np.int32)
        self._fn("get_ints")(self._h, _ip(out))
        return dict(zip(INTS, (int(v) for v in out)))

    def doubles(self):
        out = np.zeros(10, np.float64)
        self._fn("get_doubles")(self._h, _dp(out))
        return dict(zip(DOUBLES, (float(v) for v in out)))

    def array(self, name):
        i = self.ints()
        n, lam, m, mu = i["n"], i["lambda"], i["m"], i["mu"]
        size = {"xmean": n, "xold": n, "pc": n, "V": m * n, "P": m * n, "Nj": m, "Lj": m, "X": lam * n,
                "fit": lam, "prev_fit": lam, "weights": mu}[name]
        out = np.zeros(size, np.float64)
        got = self._fn("get_array")(self._h, ARRAYS[name], _dp(out))
        assert got == size, (name, got, size)
        if name in ("V", "P"):
            return out.reshape(m, n)
        if name == "X":
            return out.reshape(lam, n)
        return out

    def int_array(self, name):
        i = self.ints()
        size = i["lambda"] if name == "arindex" else i["m"]
        out = np.zeros(size, np.int32)
        self._fn("get_int_array")(self._h, INT_ARRAYS[name], _ip(out))
        return out

    def state(self):
        """Everything the device needs for teacher forcing, as numpy arrays / python scalars."""
        st = {}
        st.update(self.ints())
        st.update(self.doubles())
        for k in ("xmean", "pc", "V", "P", "Nj", "Lj", "prev_fit"):
            st[k] = self.array(k)
        for k in ("t", "vec"):
            st[k] = self.int_array(k)
        return st

    def ask(self):
        out = np.zeros(self.n, np.float64)
        self._fn("ask")(self._h, _dp(out))
        return out

    def tell(self, f):
        self._fn("tell")(self._h, float(f))

    def done(self):
        return bool(self._fn("done")(self._h))

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleLMCMA(_Base):
    """FP64 restatement; m < 1 means m = lambda (the reference's rule)."""
    _prefix = "orc_lmcma_"

    def __init__(self, n, x0=None, lam=0, m=0, lo=None, hi=None, sigma=1.0, seed=1, Z0=None):
        self._L = lib()
        self.n = n
        self._keep = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (x0, lo, hi, Z0)]
        x0, lo, hi, Z0 = self._keep
        self._h = self._L.orc_lmcma_create(n, lam, m, _dp(x0), _dp(lo), _dp(hi), float(sigma), int(seed), _dp(Z0))

    def load_state(self, st, prev_fit, Z=None):
        """Warm start from a state dict of another OracleLMCMA of the same (n, m) — see orc_lmcma_set_state."""
        a = {k: np.ascontiguousarray(st[k], np.float64) for k in ("xmean", "pc", "V", "P", "Nj", "Lj")}
        pf = np.ascontiguousarray(prev_fit, np.float64)
        t, vec = np.ascontiguousarray(st["t"], np.int32), np.ascontiguousarray(st["vec"], np.int32)
        Zc = None if Z is None else np.ascontiguousarray(Z, np.float64)
        self._L.orc_lmcma_set_state(self._h, _dp(a["xmean"]), _dp(a["pc"]), _dp(a["V"]), _dp(a["P"]), _dp(a["Nj"]), _dp(a["Lj"]),
                                    _dp(pf), _ip(t), _ip(vec), int(st["itr"]), int(st["live"]), float(st["sigma"]),
                                    float(st["s"]), _dp(Zc))

    def tell_all(self, f, Z_next=None):
        f = np.ascontiguousarray(f, np.float64)
        Z = None if Z_next is None else np.ascontiguousarray(Z_next, np.float64)
        self._L.orc_lmcma_tell_all(self._h, _dp(f), _dp(Z))


class RefLMCMA(_Base):
    """The compiled reference.  lambda < 1 -> reference default; seed must be >= 1 for determinism."""
    _prefix = "ref_lmcma_"

    def __init__(self, n, x0=None, lam=0, lo=None, hi=None, sigma=1.0, seed=1, cov=None, o0=False):
        self._L = ref(o0)
        self.n = n
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (x0, lo, hi, cov)]
        x0, lo, hi, cov = arrs
        self._h = self._L.ref_lmcma_create(_dp(x0), n, lam, _dp(lo), _dp(hi), float(sigma), _dp(cov), int(seed))

    def generation(self, problem):
        """One generation through the reference's ask/tell protocol with the cost oracle as the user
        cost (threaded across candidates).  Returns (X, f)."""
        i = self.ints()
        X = np.zeros((i["lambda"], i["n"]), np.float64)
        f = np.zeros(i["lambda"], np.float64)
        fn = C.cast(lib().orc_cost_batch_f64, C.c_void_p)
        self._L.ref_lmcma_generation(self._h, fn, C.addressof(problem.struct), _dp(X), _dp(f))
        return X, f


_BATCH_COST = C.CFUNCTYPE(None, _c_double_p, C.c_int, C.c_int, _c_double_p, C.c_void_p)
SIBLING_KINDS = {"LMCMA": 0, "SepCMA": 1, "CMAChol": 2}


class RefSibling:
    """The reference's sibling optimisers (SepCMA lmcma.hpp:152, CMAChol lmcma.hpp:211; LMCMA for a like-for-like
    run) through the shared CMABase ask / tell protocol.  CPU cross-check of solution quality only."""

    def __init__(self, kind, n, x0=None, lam=0, lo=None, hi=None, sigma=1.0, seed=1):
        self._L = ref()
        self.n = n
        x0, lo, hi = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (x0, lo, hi)]
        self._h = self._L.ref_sibling_create(SIBLING_KINDS[kind], _dp(x0), n, lam, _dp(lo), _dp(hi), float(sigma),
                                             int(seed))
        self.lam = self._L.ref_sibling_lambda(self._h)

    def run(self, generations, problem=None, func=None):
        """`generations` generations on a CostProblem (cost oracle) or on a python `func(X[k, n]) -> f[k]`.
        Returns (BestF, best_x)."""
        best_x = np.zeros(self.n, np.float64)
        if problem is not None:
            fn, ctx = C.cast(lib().orc_cost_batch_f64, C.c_void_p), C.addressof(problem.struct)
        else:
            def _cb(X, count, n, f, _ctx):
                Xa = np.ctypeslib.as_array(X, shape=(count, n))
                fa = np.ctypeslib.as_array(f, shape=(count,))
                fa[:] = func(Xa)
            keep = _BATCH_COST(_cb)
            fn, ctx = C.cast(keep, C.c_void_p), None
        best = self._L.ref_sibling_run(self._h, fn, ctx, int(generations), _dp(best_x))
        return float(best), best_x

    def close(self):
        if self._h:
            self._L.ref_sibling_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rng_gauss(seed, count, skip=0, which="oracle"):
    out = np.zeros(count, np.float64)
    if which == "oracle":
        lib().orc_rng_gauss(seed, skip, count, _dp(out))
    else:
        ref().ref_rng_gauss(seed, skip, count, _dp(out))
    return out


def rng_uniform(seed, count, which="oracle"):
    out = np.zeros(count, np.float64)
    (lib().orc_rng_uniform if which == "oracle" else ref().ref_rng_uniform)(seed, count, _dp(out))
    return out


def rank(values, which="oracle"):
    v = np.array(values, np.float64)
    ids = np.zeros(len(v), np.int32)
    if which == "oracle":
        lib().orc_rank(len(v), _dp(v), _ip(ids))
    else:
        ref().ref_myqsort(len(v), _dp(v), _ip(ids))
    return v, ids


class _Problem(C.Structure):
    _fields_ = [("dims", C.c_int), ("shape", C.c_int * 3), ("dist", _c_float_p), ("c_min", C.c_float),
                ("start", C.c_float * 3), ("goal", C.c_float * 3), ("waypoints", C.c_int),
                ("w_len", C.c_float), ("w_clr", C.c_float), ("w_col", C.c_float), ("threads", C.c_int)]


class CostProblem:
    """dist: float32 array [ny, nx] or [nz, ny, nx] (distance to nearest obstacle, 0 on obstacles)."""

    def __init__(self, dist, start, goal, waypoints, w_len=1.0, w_clr=1000.0, w_col=1e4, c_min=0.5, threads=1):
        self.dist = np.ascontiguousarray(dist, np.float32)
        dims = self.dist.ndim
        assert dims in (2, 3)
        shp = self.dist.shape[::-1]  # (nx, ny[, nz])
        s = _Problem()
        s.dims = dims
        for i in range(3):
            s.shape[i] = shp[i] if i < dims else 1
            s.start[i] = float(start[i]) if i < dims else 0.0
            s.goal[i] = float(goal[i]) if i < dims else 0.0
        s.dist = _fp(self.dist)
        s.c_min = c_min
        s.waypoints = waypoints
        s.w_len, s.w_clr, s.w_col = w_len, w_clr, w_col
        s.threads = threads
        self.struct = s
        self.dims, self.waypoints, self.n = dims, waypoints, dims * waypoints

    def evaluate(self, X):
        X = np.ascontiguousarray(X, np.float32).reshape(-1, self.n)
        cnt = X.shape[0]
        f = np.zeros(cnt, np.float64)
        ncoll = np.zeros(cnt, np.int32)
        nsamp = np.zeros(cnt, np.int32)
        ln = np.zeros(cnt, np.float64)
        clr = np.zeros(cnt, np.float64)
        lib().orc_cost_batch(C.addressof(self.struct), _fp(X), cnt, _dp(f), _ip(ncoll), _ip(nsamp), _dp(ln), _dp(clr))
        return {"f": f, "ncoll": ncoll, "nsamp": nsamp, "length": ln, "clearance": clr}

    def sample(self, Q):
        """Per-sample pieces for points Q [k, dims] (FP32): (linear cell or -1, collides, state cost)."""
        Q = np.ascontiguousarray(Q, np.float32).reshape(-1, self.dims)
        q3 = np.zeros((len(Q), 3), np.float32)
        q3[:, :self.dims] = Q
        cell = np.zeros(len(Q), np.int64)
        hit = np.zeros(len(Q), np.int32)
        g = np.zeros(len(Q), np.float64)
        lib().orc_cost_sample(C.addressof(self.struct), _fp(q3), len(Q), cell.ctypes.data_as(_c_i64_p), _ip(hit), _dp(g))
        return cell, hit, g

    def trace(self, x, max_cells=1 << 22):
        x = np.ascontiguousarray(x, np.float32)
        cells = np.zeros(max_cells, np.int64)
        ns = lib().orc_cost_trace(C.addressof(self.struct), _fp(x), cells.ctypes.data_as(_c_i64_p), max_cells)
        return cells[:min(ns, max_cells)].copy()


def edt_exact(occ, clamp=0.0):
    occ = np.ascontiguousarray(occ, np.uint8)
    shp = np.array(occ.shape[::-1], np.int32)
    out = np.zeros(occ.shape, np.float32)
    lib().orc_edt_exact(occ.ctypes.data_as(C.c_void_p), occ.ndim, _ip(shp), float(clamp), _fp(out))
    return out


def ssedt8_sq(seed_mask):
    m = np.ascontiguousarray(seed_mask, np.uint8)
    out = np.zeros(m.shape, np.int32)
    lib().orc_8ssedt_sq(m.ctypes.data_as(C.c_void_p), m.shape[1], m.shape[0], _ip(out))
    return out


def ssedt8_signed(occ):
    m = np.ascontiguousarray(occ, np.uint8)
    out = np.zeros(m.shape, np.int32)
    lib().orc_8ssedt_signed(m.ctypes.data_as(C.c_void_p), m.shape[1], m.shape[0], _ip(out))
    return out
