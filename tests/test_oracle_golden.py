"""CPU: the C++ oracle against the committed golden fixtures (made by tests/golden/
make_golden.py from the cv2-based restatement oracle/py_ref.py)."""
import zlib

import numpy as np
import pytest

from conftest import KP_FIELDS, load_golden


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


@pytest.mark.parametrize("name", ["orb_320x240_corridor.npz", "orb_640x480_room.npz"])
def test_orb_oracle_matches_golden(orc, name):
    g = load_golden(name)
    o = orc.OrbOracle(int(g["nfeatures"]), 1.2, 8, 20, 7)
    assert o.features_per_level() == list(g["per_level"])
    assert o.umax() == list(g["umax"])
    assert np.array_equal(np.array(o.scale_factors(), np.float32), g["scale"])
    kps, desc = o.extract(g["gray"])
    for l in range(8):
        assert crc(o.level(l, bordered=True)) == g["pyr_crc_%d" % l], "pyramid level %d" % l
        c = o.candidates(l)
        assert np.array_equal(c.astype(np.int16), g["cands_%d" % l]), "FAST candidates level %d (order included)" % l
        lk = o.level_keypoints(l)
        gl = g["lkp_%d" % l]
        assert len(lk) == len(gl)
        assert np.array_equal(np.stack([lk["x"], lk["y"], lk["response"], lk["angle"]], 1), gl) if len(gl) else True
        if len(gl):
            assert crc(o.blurred(l)) == g["blur_crc_%d" % l], "blur level %d" % l
    assert np.array_equal(o.level(3, bordered=True), g["pyr_level3"])
    assert len(kps) == len(g["kps"])
    for f in KP_FIELDS:
        assert np.array_equal(kps[f], g["kps"][f]), f
    assert np.array_equal(desc, g["desc"])


def test_glibc_rand_stream_matches_libc(orc):
    """the declared rand() stream of CylinderSeg (App. B.9) is glibc's, checked against the libc of this box"""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 20260000):
        libc.srand(seed)
        ref = np.array([libc.rand() for _ in range(1000)], np.int32)
        assert np.array_equal(orc.glibc_rand(seed, 1000), ref)
    libc.srand(1)
    assert orc.glibc_rand(1, 2).tolist() == [1804289383, 846930886]


def test_features_per_level_reference_values(orc):
    # ORBextractor(1000, 1.2, 8, ...) -> mnFeaturesPerLevel
    assert orc.OrbOracle(1000).features_per_level() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert orc.OrbOracle(1000).umax() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


@pytest.mark.parametrize("name", ["cape_640x480_corridor_m.npz", "cape_640x480_room_mm.npz",
                                  "cape_320x240_room_mm_cell10.npz", "cape_640x480_pillars_mm_cyl.npz",
                                  "cape_640x480_pillars_m_cyl.npz"])
def test_cape_oracle_matches_golden(orc, name):
    g = load_golden(name)
    depth = g["depth_q"].astype(np.float32) * np.float32(1.0 / 5000.0) * g["unit"]
    h, w = depth.shape
    cell = int(g["cell"])
    cyl_on = "cylinder" in g
    o = orc.CapeOracle(h, w, cell, cell, cyl_on, float(g["min_cos"]), float(g["max_merge"]))
    cloud = o.depth_to_cloud(depth, *[float(v) for v in g["K"]])
    assert crc(cloud) == g["cloud_crc"]
    if cyl_on:
        seg, planes, ncyl_final, cyls = o.process_full(cloud)
        assert ncyl_final == int(g["nr_cylinders_final"]) and len(cyls) == len(g["cyl_radius"]) > 0
        cm, ce = o.cyl_maps()
        assert np.array_equal(cm, g["cyl_map"]) and np.array_equal(ce, g["cyl_eroded"])
        assert np.allclose(cyls["radius"], g["cyl_radius"], rtol=1e-6)
        assert np.allclose(cyls["center"], g["cyl_center"], rtol=1e-7, atol=1e-7)
        # the PCA axis sign is solver-defined (LAPACK there, Jacobi here)
        assert np.allclose(np.abs(cyls["axis"]), np.abs(g["cyl_axis"]), atol=1e-9)
        assert int(seg.max()) > 50, "a cylinder label (50 + k) is painted"
    else:
        seg, planes = o.process(cloud)
    cells = o.cells()
    assert np.array_equal(cells["planar"].astype(np.uint8), g["cell_planar"])
    assert np.array_equal(cells["nr_pts"], g["cell_nr_pts"])
    sums = np.stack([cells[f] for f in ("x_acc", "y_acc", "z_acc", "xx_acc", "yy_acc", "zz_acc", "xy_acc", "xz_acc", "yz_acc")], 1)
    assert np.array_equal(sums, g["cell_sums"]), "declared float summation tree (App. B.1)"
    fitted = g["cell_mse"] != 0
    # the golden eigen-solve is LAPACK (numpy eigh), the oracle's is Jacobi: same to ~1e-12
    assert np.allclose(cells["normal"][fitted], g["cell_normal"][fitted], atol=1e-9, rtol=0)
    assert np.allclose(cells["d"][fitted], g["cell_d"][fitted], rtol=1e-9, atol=1e-9)
    pm, em = o.grid_maps()
    assert np.array_equal(pm, g["plane_map"]) and np.array_equal(em, g["eroded_map"])
    assert np.array_equal(seg, g["seg"])
    assert len(planes) == len(g["plane_d"])
    assert np.array_equal(planes["nr_pts"], g["plane_nr_pts"])
    assert np.allclose(planes["normal"], g["plane_normal"], atol=1e-9, rtol=0)
    assert np.allclose(planes["d"], g["plane_d"], rtol=1e-9)
    assert np.allclose(planes["MSE"], g["plane_mse"], rtol=1e-5)
    assert np.allclose(planes["score"], g["plane_score"], rtol=1e-4)
