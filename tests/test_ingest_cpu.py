"""Map ingest formats: host-side parsers of the C ABI against files written here and, where
/root/reference exists, against the reference's own bundled files."""
import os
import re
import struct

import numpy as np
import pytest

from conftest import REFERENCE_PRESENT
from lmcma_path_planner_b200 import maps
from lmcma_path_planner_b200 import _capi as K

REF = "/root/reference/sample_based_optimisation_based_path_planner"


def _write_bmp24(path, rgb, bottom_up=True):
    h, w, _ = rgb.shape
    stride = (w * 3 + 3) & ~3
    rows = []
    order = range(h - 1, -1, -1) if bottom_up else range(h)
    for y in order:
        line = rgb[y, :, ::-1].astype(np.uint8).tobytes()
        rows.append(line + b"\0" * (stride - len(line)))
    data = b"".join(rows)
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(data), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, w, h if bottom_up else -h, 1, 24, 0,
                                                                               len(data), 2835, 2835, 0, 0)
    open(path, "wb").write(hdr + data)


def test_bmp_loader_applies_the_reference_rule(tmp_path):
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, size=(7, 5, 3), dtype=np.uint8)          # width 5: rows are padded to 16 bytes
    rgb[0, 0] = (255, 127, 255); rgb[0, 1] = (0, 128, 0)                # g = 127 -> obstacle, g = 128 -> free
    want = (rgb[:, :, 1] < 128).astype(np.uint8)
    for bottom_up in (True, False):
        p = str(tmp_path / ("a%d.bmp" % bottom_up))
        _write_bmp24(p, rgb, bottom_up)
        assert np.array_equal(maps.load_bmp(p), want)
    with pytest.raises(K.LmcmaError):
        maps.load_bmp(str(tmp_path / "missing.bmp"))


def test_binvox_loader_run_lengths_and_axis_order(tmp_path):
    d, h, w = 3, 4, 5                                                    # dim line: depth height width
    rng = np.random.default_rng(1)
    vox = (rng.random(d * h * w) < 0.4).astype(np.uint8)                 # file order: i -> y = i % w, z = (i / w) % h, x = i / (w h)
    runs = bytearray()
    i = 0
    while i < len(vox):
        j = i
        while j < len(vox) and vox[j] == vox[i] and j - i < 255:
            j += 1
        runs += bytes([int(vox[i]), j - i])
        i = j
    p = str(tmp_path / "t.binvox")
    open(p, "wb").write(b"#binvox 1\ndim %d %d %d\ntranslate -0.5 0.25 1\nscale 2.5\ndata\n" % (d, h, w) + bytes(runs))
    occ, tr, sc = maps.load_binvox(p)
    assert occ.shape == (h, w, d) and tr.tolist() == [-0.5, 0.25, 1.0] and sc == 2.5      # [nz, ny, nx] = [height, width, depth]
    for i in range(len(vox)):
        y, z, x = i % w, (i // w) % h, i // (w * h)
        assert occ[z, y, x] == vox[i]


def test_text_matrix_loader(tmp_path):
    m = np.arange(12, dtype=np.float64).reshape(3, 4) / 7.0
    p = str(tmp_path / "mat.txt")
    open(p, "w").write("\n".join(",".join(repr(float(v)) for v in row) for row in m) + "\n")
    assert np.array_equal(maps.load_text_matrix(p), m)
    open(p, "w").write("1,2,3\n4,5\n")
    with pytest.raises(K.LmcmaError):
        maps.load_text_matrix(p)


def _write_bt(path, occ_keys, free_keys, res, depth=16):
    """An OctoMap binary tree written here from sets of leaf keys: 2 bits per child (01 occupied, 10 free, 11 inner, 00
    unknown; bit 2i first), inner nodes depth first, a node whose 8 children are equal leaves pruned into one leaf."""
    def build(keys_occ, keys_free, level):                               # -> 'occ' | 'free' | None | [8 children]
        if level == depth:
            return "occ" if keys_occ else ("free" if keys_free else None)
        if not keys_occ and not keys_free:
            return None
        bit = depth - 1 - level
        kids = []
        for i in range(8):
            sel = lambda ks: [k for k in ks if ((k[0] >> bit) & 1, (k[1] >> bit) & 1, (k[2] >> bit) & 1) == (i & 1, (i >> 1) & 1, (i >> 2) & 1)]
            kids.append(build(sel(keys_occ), sel(keys_free), level + 1))
        if all(k == "occ" for k in kids):
            return "occ"
        if all(k == "free" for k in kids):
            return "free"
        return kids
    out = bytearray()
    count = [1]                                                          # the root
    def emit(node):
        bits = 0
        for i, k in enumerate(node):
            code = 0 if k is None else (2 if k == "occ" else (1 if k == "free" else 3))    # value of (bit 2i) | (bit 2i+1) << 1
            bits |= code << (2 * i)
            count[0] += k is not None
        out.extend(struct.pack("<H", bits))
        for k in node:
            if isinstance(k, list):
                emit(k)
    root = build(list(occ_keys), list(free_keys), 0)
    assert isinstance(root, list)
    emit(root)
    open(path, "wb").write(b"# Octomap OcTree binary file\n# comment\nid OcTree\nsize %d\nres %s\ndata\n" % (count[0], repr(res).encode())
                           + bytes(out))


def test_bt_loader_leaves_pruned_cubes_and_header(tmp_path):
    rng = np.random.default_rng(3)
    base = np.array([32768 - 5, 32768 + 8, 32768 + 16])                  # straddles the sign boundary of the x axis
    cells = rng.random((11, 9, 14)) < 0.3                                # [z, y, x]
    cells[0:4, 0:4, 8:12] = True                                         # key-aligned 4^3 cube: pruned two levels up
    zz, yy, xx = np.nonzero(cells)
    occ_keys = {(int(base[0] + x), int(base[1] + y), int(base[2] + z)) for x, y, z in zip(xx, yy, zz)}
    free_keys = {(int(base[0] + x), int(base[1] + y), int(base[2] + 30)) for x in range(14) for y in range(9)} - occ_keys
    p = str(tmp_path / "t.bt")
    _write_bt(p, occ_keys, free_keys, 0.05)
    occ, org, res = maps.load_bt(p)
    lo = np.array([min(k[c] for k in occ_keys) for c in range(3)])
    hi = np.array([max(k[c] for k in occ_keys) for c in range(3)]) + 1
    assert res == 0.05 and org.tolist() == lo.tolist() and occ.shape == tuple((hi - lo)[::-1])
    want = np.zeros(occ.shape, np.uint8)
    for k in occ_keys:
        want[k[2] - lo[2], k[1] - lo[1], k[0] - lo[0]] = 1
    assert np.array_equal(occ, want)                                     # free leaves and unknown space are both 0
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-3])                                        # truncated data
    with pytest.raises(K.LmcmaError):
        maps.load_bt(p)
    open(p, "wb").write(raw.replace(b"id OcTree", b"id ColorOcTree"))    # other tree types are not binary-compatible
    with pytest.raises(K.LmcmaError):
        maps.load_bt(p)
    open(p, "wb").write(re.sub(rb"size \d+", b"size 7", raw))             # node count does not match
    with pytest.raises(K.LmcmaError):
        maps.load_bt(p)


@pytest.mark.skipif(not REFERENCE_PRESENT, reason="needs /root/reference (authoring container)")
def test_reference_files(golden_maps):
    for name in ("problem1", "problem2"):
        assert np.array_equal(maps.load_bmp(os.path.join(REF, "images", name + ".bmp")), golden_maps[name])
    occ, tr, sc = maps.load_binvox(os.path.join(REF, "files", "mesh_files", "Dude.binvox"))
    assert occ.shape == (256, 256, 256)
    assert abs(occ.mean() - 0.0174) < 5e-4                               # 1.74 % occupied


@pytest.mark.skipif(not REFERENCE_PRESENT, reason="needs /root/reference (authoring container)")
@pytest.mark.parametrize("name", ["Dude", "room"])
def test_reference_bt_equals_its_binvox_source(name):
    """The bundled .bt files were made from the bundled .binvox files by binvox2bt.cpp:250-271: voxel (x, y, z) -> point
    (float)(v res + t + 1e-6) -> key floor(p / res) + 32768, res = scale / depth.  Both loaders must agree cell for cell."""
    d = os.path.join(REF, "files", "mesh_files")
    occ, org, res = maps.load_bt(os.path.join(d, name + ".binvox.bt"))
    vox, tr, sc = maps.load_binvox(os.path.join(d, name + ".binvox"))
    assert abs(res - sc / vox.shape[2]) < 1e-6 * res                     # the header prints 5-6 significant digits
    zz, yy, xx = np.nonzero(vox)
    r = sc / vox.shape[2]
    key = lambda v, t: np.floor((v * r + t + 0.000001).astype(np.float32).astype(np.float64) / res).astype(np.int64) + 32768
    ix, iy, iz = key(xx, tr[0]) - org[0], key(yy, tr[1]) - org[1], key(zz, tr[2]) - org[2]
    want = np.zeros_like(occ)
    want[iz, iy, ix] = 1
    assert np.array_equal(occ, want) and occ.sum() == vox.sum()


def test_corrupted_files_are_rejected_or_parsed_never_fatal(tmp_path):
    """Seeded mutations (truncation, byte flips, extreme 32-bit header fields, trailing garbage) of one valid file per
    format: every call returns an occupancy grid or raises LmcmaError - the parsers bound every size they read."""
    rng = np.random.default_rng(7)
    seeds = {}
    p = str(tmp_path / "s.bmp")
    _write_bmp24(p, rng.integers(0, 256, size=(9, 6, 3), dtype=np.uint8))
    seeds["bmp"] = (p, maps.load_bmp)
    p = str(tmp_path / "s.binvox")
    open(p, "wb").write(b"#binvox 1\ndim 4 4 4\ntranslate 0 0 0\nscale 1\ndata\n" + bytes([0, 30, 1, 34]))
    seeds["binvox"] = (p, maps.load_binvox)
    p = str(tmp_path / "s.txt")
    open(p, "w").write("1 2 3\n4 5 6\n")
    seeds["txt"] = (p, maps.load_text_matrix)
    p = str(tmp_path / "s.bt")
    _write_bt(p, {(32768 + x, 32768 + y, 32768) for x in range(5) for y in range(3)}, {(32768, 32768, 32770)}, 0.1)
    seeds["bt"] = (p, maps.load_bt)
    for kind, (path, load) in seeds.items():
        load(path)                                                       # the seed itself is valid
        data = open(path, "rb").read()
        outcomes = set()
        for it in range(80):
            b = bytearray(data)
            mode = it % 4
            if mode == 0:
                b = b[:int(rng.integers(0, len(b)))]
            elif mode == 1:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            elif mode == 2:
                pos = int(rng.integers(0, max(1, min(len(b), 64) - 4)))
                b[pos:pos + 4] = struct.pack("<i", int(rng.choice([-1, 0, 2 ** 31 - 1, -2 ** 31, 65536])))
            else:
                b += bytes(rng.integers(0, 256, 16, dtype=np.uint8))
            q = str(tmp_path / ("m." + kind))
            open(q, "wb").write(bytes(b))
            try:
                load(q)
                outcomes.add("parsed")
            except K.LmcmaError:
                outcomes.add("rejected")
        assert "rejected" in outcomes, kind
