"""Synthetic maps and queries of the BASELINE.json shapes plus the reference's
bundled 2-D maps.  Host-side workload generation only (numpy / scipy): nothing here is on the hot path.

Conventions (planner.cpp:597-602): 2-D maps are [row = y][col = x]; 3-D maps are [z][y][x]; a
distance map holds the Euclidean distance (cells) to the nearest obstacle and 0 on obstacles."""
import numpy as np


def distance_field(occ, clamp=0.0):
    """Exact EDT of an occupancy grid (non-zero = obstacle) as float32, optionally clamped."""
    from scipy import ndimage
    d = ndimage.distance_transform_edt(np.asarray(occ) == 0).astype(np.float32)
    if clamp > 0:
        np.minimum(d, np.float32(clamp), out=d)
    return d


def two_bars_occupancy():
    """The reference's hard-coded 100x100 map (planner.cpp:269-297), as [row=y][col=x]... the reference writes
    EDT_Matrix(x, y) = 1 for x in [0,60), y in [65,75) and x in [40,99), y in [35,45)."""
    occ = np.zeros((100, 100), np.uint8)
    occ[65:75, 0:60] = 1
    occ[35:45, 40:99] = 1
    return occ


def problem_occupancy(which):
    """problem1.bmp / problem2.bmp of the reference as rectangles (verified against the
    BMPs by tests/test_golden_maps.py): black (g < 128, planner.cpp:515) = obstacle."""
    occ = np.zeros((100, 100), np.uint8)
    if which == 1:
        occ[0:47, 23:37] = 1
        occ[46:100, 60:77] = 1
    elif which == 2:
        occ[0:46, 62:75] = 1
        occ[50:100, 24:39] = 1
    else:
        raise ValueError(which)
    return occ


def random_boxes_occupancy(shape, n_boxes, side_lo, side_hi, seed, border=2, clear=()):
    """Axis-aligned random boxes + a solid border; `clear` = [(centre_xyz, radius)] balls kept free."""
    rng = np.random.default_rng(seed)
    dims = len(shape)
    occ = np.zeros(shape, np.uint8)
    sides = rng.integers(side_lo, side_hi + 1, size=(n_boxes, dims))
    for i in range(n_boxes):
        lo = [int(rng.integers(0, max(1, shape[a] - sides[i, a]))) for a in range(dims)]
        sl = tuple(slice(lo[a], lo[a] + int(sides[i, a])) for a in range(dims))
        occ[sl] = 1
    if border > 0:
        for a in range(dims):
            sl = [slice(None)] * dims
            sl[a] = slice(0, border); occ[tuple(sl)] = 1
            sl[a] = slice(shape[a] - border, shape[a]); occ[tuple(sl)] = 1
    for centre, radius in clear:
        c = np.asarray(centre)[::-1]   # xyz -> array index order
        lo = np.maximum(0, np.floor(c - radius).astype(int))
        hi = np.minimum(np.array(shape), np.ceil(c + radius).astype(int) + 1)
        grids = np.ogrid[tuple(slice(lo[a], hi[a]) for a in range(dims))]
        d2 = sum((grids[a] - c[a]) ** 2 for a in range(dims))
        sub = occ[tuple(slice(lo[a], hi[a]) for a in range(dims))]
        sub[d2 <= radius * radius] = 0
    return occ


def config2_map(size=4096, n_rects=2048, seed=42, clamp=256.0):
    """C2: size^2 occupancy, border of 2 cells + n_rects rectangles of side U[8,128], start/goal discs of
    radius 64 cleared; f32 EDT clamped at 256 cells.  Returns (dist, start_xy, goal_xy)."""
    start, goal = (64.0, 64.0), (size - 64.0, size - 64.0)
    scale = size / 4096.0
    occ = random_boxes_occupancy((size, size), max(1, int(n_rects * scale * scale)), 8, 128, seed, border=2,
                                 clear=[(start, 64.0 * min(1.0, scale * 4)), (goal, 64.0 * min(1.0, scale * 4))])
    return distance_field(occ, clamp), start, goal


def config4_map(size=512, n_boxes=4096, seed=43, clamp=64.0):
    """C4: size^3 occupancy with n_boxes boxes of side U[4,48]; f32 EDT clamped at 64."""
    start, goal = (16.0, 16.0, 16.0), (size - 16.0, size - 16.0, size - 16.0)
    scale = size / 512.0
    occ = random_boxes_occupancy((size, size, size), max(1, int(n_boxes * scale ** 3)), 4, 48, seed, border=0,
                                 clear=[(start, 12.0), (goal, 12.0)])
    return distance_field(occ, clamp), start, goal


def straight_line(start, goal, waypoints):
    """Dimension-major x0 (x[d*W + w], lmcma.cpp:786-791): W interior points evenly spaced on start->goal."""
    s, g = np.asarray(start, np.float64), np.asarray(goal, np.float64)
    t = (np.arange(1, waypoints + 1) / (waypoints + 1.0))[None, :]
    return (s[:, None] + (g - s)[:, None] * t).reshape(-1)


def box_bounds(shape_xyz, waypoints):
    """lo / hi of the state space, per parameter (planner.cpp:696-697: [0, size-1] per axis)."""
    dims = len(shape_xyz)
    lo = np.zeros(dims * waypoints)
    hi = np.repeat(np.asarray(shape_xyz, np.float64) - 1.0, waypoints)
    return lo, hi


def random_queries(dist, count, seed, min_sep):
    """C3: start / goal pairs drawn uniformly from free cells with |goal - start|_2 >= min_sep (xy[z] order)."""
    rng = np.random.default_rng(seed)
    free = np.argwhere(dist > 1.5)
    starts, goals = [], []
    while len(starts) < count:
        a = free[rng.integers(0, len(free), size=count)]
        b = free[rng.integers(0, len(free), size=count)]
        ok = np.linalg.norm((a - b).astype(np.float64), axis=1) >= min_sep
        for i in np.nonzero(ok)[0]:
            if len(starts) < count:
                starts.append(a[i][::-1].astype(np.float32))
                goals.append(b[i][::-1].astype(np.float32))
    return np.array(starts), np.array(goals)


# ---- the library's map ingest / distance transform entry points (include/lmcma_b200.h) ----
def load_bmp(path):
    """Occupancy (1 = obstacle, the reference's g < 128 rule) of a 24/32-bit BMP as uint8 [height, width]."""
    import ctypes as C
    from . import _capi as K
    w, h = C.c_int32(0), C.c_int32(0)
    K.check(K.lib().lmcma_b200_load_bmp(path.encode(), None, 0, C.byref(w), C.byref(h)))
    occ = np.zeros((h.value, w.value), np.uint8)
    K.check(K.lib().lmcma_b200_load_bmp(path.encode(), occ.ctypes.data_as(C.POINTER(C.c_uint8)), occ.size, C.byref(w), C.byref(h)))
    return occ


def load_binvox(path):
    """Dense voxel occupancy [nz, ny, nx] (uint8) + header (translate, scale) of a binvox file."""
    import ctypes as C
    from . import _capi as K
    shp = np.zeros(3, np.int32)
    tr = np.zeros(3, np.float64)
    sc = C.c_double(0)
    K.check(K.lib().lmcma_b200_load_binvox(path.encode(), None, 0, K.iptr(shp), K.dptr(tr), C.byref(sc)))
    occ = np.zeros((int(shp[2]), int(shp[1]), int(shp[0])), np.uint8)
    K.check(K.lib().lmcma_b200_load_binvox(path.encode(), occ.ctypes.data_as(C.POINTER(C.c_uint8)), occ.size, K.iptr(shp), K.dptr(tr),
                                         C.byref(sc)))
    return occ, tr, sc.value


def load_bt(path):
    """Dense occupancy [nz, ny, nx] (uint8, 1 = occupied) of the bounding box of the occupied leaves of an OctoMap binary
    tree (.bt), the key of its first cell per axis (x, y, z) and the leaf size."""
    import ctypes as C
    from . import _capi as K
    shp = np.zeros(3, np.int32)
    org = np.zeros(3, np.int32)
    res = C.c_double(0)
    K.check(K.lib().lmcma_b200_load_bt(path.encode(), None, 0, K.iptr(shp), K.iptr(org), C.byref(res)))
    occ = np.zeros((int(shp[2]), int(shp[1]), int(shp[0])), np.uint8)
    K.check(K.lib().lmcma_b200_load_bt(path.encode(), occ.ctypes.data_as(C.POINTER(C.c_uint8)), occ.size, K.iptr(shp), K.iptr(org),
                                     C.byref(res)))
    return occ, org, res.value


def load_text_matrix(path):
    import ctypes as C
    from . import _capi as K
    r, c = C.c_int32(0), C.c_int32(0)
    K.check(K.lib().lmcma_b200_load_text_matrix(path.encode(), None, 0, C.byref(r), C.byref(c)))
    out = np.zeros((r.value, c.value), np.float64)
    K.check(K.lib().lmcma_b200_load_text_matrix(path.encode(), K.dptr(out), out.size, C.byref(r), C.byref(c)))
    return out


def edt_device(occ, clamp=0.0, device=0):
    """Exact Euclidean distance transform on the GPU (k_edt.cuh): occupancy (non-zero = obstacle) -> float32 distances."""
    import ctypes as C
    from . import _capi as K
    o = np.ascontiguousarray(np.asarray(occ) != 0, np.uint8)
    shp = np.array(o.shape[::-1], np.int32)
    out = np.zeros(o.shape, np.float32)
    K.check(K.lib().lmcma_b200_edt(device, o.ndim, K.iptr(shp), o.ctypes.data_as(C.POINTER(C.c_uint8)), float(clamp), K.fptr(out)))
    return out
