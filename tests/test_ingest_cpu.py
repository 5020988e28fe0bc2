"""Map ingest formats: host-side parsers of the C ABI against files written here and, where
/root/reference exists, against the reference's own bundled files."""
import os
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


@pytest.mark.skipif(not REFERENCE_PRESENT, reason="needs /root/reference (authoring container)")
def test_reference_files(golden_maps):
    for name in ("problem1", "problem2"):
        assert np.array_equal(maps.load_bmp(os.path.join(REF, "images", name + ".bmp")), golden_maps[name])
    occ, tr, sc = maps.load_binvox(os.path.join(REF, "files", "mesh_files", "Dude.binvox"))
    assert occ.shape == (256, 256, 256)
    assert abs(occ.mean() - 0.0174) < 5e-4                               # 1.74 % occupied
