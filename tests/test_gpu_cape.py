"""GPU: the CUDA CAPE path through the C ABI against the CPU oracle and golden fixtures.
Bars (BASELINE.json north_star): cell segmentation and seg_output bit-exact, plane normals
and offsets within 1e-5 (here they are bit-identical to the oracle, which runs the same
Jacobi eigen-solve; the golden fixtures use LAPACK, hence the 1e-9 tolerance there)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
MC = float(np.float32(np.cos(np.pi / 12)))
TOL = 1e-5


def compare_with_oracle(cp, o, depth, K, f=0, exact=True):
    cloud = o.depth_to_cloud(depth, *K)
    oseg, oplanes = o.process(cloud)
    assert np.array_equal(cp.cloud(f), cloud), "organized cloud"
    cg, co = cp.cells(f), o.cells()
    for n in ("nr_pts", "planar", "x_acc", "y_acc", "z_acc", "xx_acc", "yy_acc", "zz_acc", "xy_acc", "xz_acc", "yz_acc"):
        assert np.array_equal(cg[n], co[n]), "cell %s" % n
    assert np.allclose(cg["normal"], co["normal"], atol=TOL, rtol=0) and np.allclose(cg["d"], co["d"], atol=TOL, rtol=1e-9)
    if exact:
        for n in ("normal", "d", "mean", "MSE", "score"):
            assert np.array_equal(cg[n], co[n]), "cell %s (same Jacobi sequence)" % n
    pm, em = cp.grid_maps(f)
    opm, oem = o.grid_maps()
    assert np.array_equal(pm, opm), "grid_plane_seg_map"
    assert np.array_equal(em, oem), "eroded map"
    return oseg, oplanes


def same_planes(a, b):
    """field-wise bit equality (the struct has 4 padding bytes whose content is unspecified)"""
    return len(a) == len(b) and all(np.array_equal(a[n], b[n]) for n in a.dtype.names)


def check_planes(planes, oplanes):
    assert len(planes) == len(oplanes)
    assert np.array_equal(planes["nr_pts"], oplanes["nr_pts"])
    assert np.allclose(planes["normal"], oplanes["normal"], atol=TOL, rtol=0)
    assert np.allclose(planes["d"], oplanes["d"], atol=TOL, rtol=1e-9)
    assert np.allclose(planes["MSE"], oplanes["MSE"], rtol=1e-5) and np.allclose(planes["score"], oplanes["score"], rtol=1e-4)


@pytest.mark.parametrize("w,h,scene,seed,unit,cell,mmd", [
    (640, 480, 0, 20260000, 1.0, 20, 50.0),        # DR-SLAM feeds metres (Frame.cc:113-115)
    (640, 480, 1, 20260077, 1.0, 20, 50.0),
    (640, 480, 0, 20260100, 1000.0, 20, 50.0),     # CAPE's native millimetres
    (640, 480, 2, 20260100, 1000.0, 20, 900.0),    # reference default max_merge_dist
    (640, 480, 1, 20260012, 1000.0, 10, 50.0),     # Realsense.yaml PATCH_SIZE 10 (100-point cells: scalar tail path)
    (1280, 720, 2, 20260140, 1000.0, 20, 50.0),    # configs[4] geometry
])
def test_cape_parity(drfe, orc, w, h, scene, seed, unit, cell, mmd):
    _, depth, K = drfe.synth_frame(w, h, scene, seed, unit)
    cp = drfe.CAPE(h, w, cell, cell, False, MC, mmd)
    npl, ncyl, seg, planes, cyl = cp.process_depth(depth, *K)
    o = orc.CapeOracle(h, w, cell, cell, False, MC, mmd)
    oseg, oplanes = compare_with_oracle(cp, o, depth, K)
    assert ncyl == 0 and npl == len(oplanes)
    assert np.array_equal(seg, oseg), "seg_output: %d px differ" % (seg != oseg).sum()
    check_planes(planes, oplanes)
    cp.close()


def test_size_not_a_multiple_of_the_cell(drfe, orc):
    """CAPE truncates to whole cells (`w / cell`, CAPE.cpp:18-19): the pixels right of / below the last full cell are
    never labelled, everything else is as on the cropped image."""
    _, depth, K = drfe.synth_frame(656, 490, 1, 20260005, 1000.0)          # 32 x 24 full cells + a 16 x 10 px margin
    cp = drfe.CAPE(490, 656, 20, 20, False, MC, 50.0)
    npl, _, seg, planes, _ = cp.process_depth(depth, *K)
    o = orc.CapeOracle(490, 656, 20, 20, False, MC, 50.0)
    oseg, oplanes = o.process(o.depth_to_cloud(depth, *K))
    assert npl == len(oplanes) > 0 and np.array_equal(seg, oseg)
    assert not seg[480:].any() and not seg[:, 640:].any()
    check_planes(planes, oplanes)
    # batch of 3 with the paint / margin path, against the single-frame results
    cp3 = drfe.CAPE(490, 656, 20, 20, False, MC, 50.0, max_batch=3)
    d3 = np.stack([depth, depth[::-1].copy(), depth])
    cp3.enqueue_depth(d3, *K, nframes=3)
    seg3 = cp3.download()[0]
    assert np.array_equal(seg3[0], seg) and np.array_equal(seg3[2], seg)


def test_process_cloud_entry_equals_depth_entry(drfe, orc):
    """CAPE::process(cloud_array, ...) boundary: same result as the fused depth entry."""
    _, depth, K = drfe.synth_frame(640, 480, 1, 20260055, 1000.0)
    o = orc.CapeOracle(480, 640, 20, 20, False, MC, 50.0)
    cloud = o.depth_to_cloud(depth, *K)
    cp = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0)
    a = cp.process(cloud)
    b = cp.process_depth(depth, *K)
    assert a[0] == b[0] and np.array_equal(a[2], b[2]) and same_planes(a[3], b[3])
    oseg, oplanes = o.process(cloud)
    assert np.array_equal(a[2], oseg)
    pre = np.full((480, 640), 9, np.uint8)                      # reference only writes labelled pixels
    out = cp.process(cloud, seg_output=pre)[2]
    assert np.array_equal(out[oseg > 0], oseg[oseg > 0]) and np.all(out[oseg == 0] == 9)


@pytest.mark.parametrize("name", ["cape_640x480_corridor_m.npz", "cape_640x480_room_mm.npz",
                                  "cape_320x240_room_mm_cell10.npz"])
def test_cape_matches_golden(drfe, name):
    g = load_golden(name)
    depth = g["depth_q"].astype(np.float32) * np.float32(1.0 / 5000.0) * g["unit"]
    h, w = depth.shape
    cell = int(g["cell"])
    cp = drfe.CAPE(h, w, cell, cell, False, float(g["min_cos"]), float(g["max_merge"]))
    npl, _, seg, planes, _ = cp.process_depth(depth, *[float(v) for v in g["K"]])
    cells = cp.cells()
    assert np.array_equal(cells["planar"].astype(np.uint8), g["cell_planar"])
    sums = np.stack([cells[f] for f in ("x_acc", "y_acc", "z_acc", "xx_acc", "yy_acc", "zz_acc", "xy_acc", "xz_acc", "yz_acc")], 1)
    assert np.array_equal(sums, g["cell_sums"])
    pm, em = cp.grid_maps()
    assert np.array_equal(pm, g["plane_map"]) and np.array_equal(em, g["eroded_map"])
    assert np.array_equal(seg, g["seg"])
    assert npl == len(g["plane_d"])
    assert np.allclose(planes["normal"], g["plane_normal"], atol=TOL, rtol=0)
    assert np.allclose(planes["d"], g["plane_d"], atol=TOL, rtol=1e-9)


def test_empty_depth_gives_no_planes(drfe, orc):
    cp = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0)
    npl, ncyl, seg, planes, _ = cp.process_depth(np.zeros((480, 640), np.float32), 525.0, 525.0, 319.5, 239.5)
    assert npl == 0 and ncyl == 0 and not seg.any() and len(planes) == 0
    assert not cp.cells()["planar"].any()


def test_single_wall_one_plane(drfe, orc):
    """A fronto-parallel wall in millimetres: one plane, normal (0,0,-1), d = depth."""
    depth = np.full((480, 640), 2000.0, np.float32)
    depth += np.random.default_rng(3).normal(0, 2.0, depth.shape).astype(np.float32)
    K = (525.0, 525.0, 319.5, 239.5)
    cp = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0)
    npl, _, seg, planes, _ = cp.process_depth(depth, *K)
    o = orc.CapeOracle(480, 640, 20, 20, False, MC, 50.0)
    oseg, oplanes = compare_with_oracle(cp, o, depth, K)
    assert npl == 1 and np.array_equal(seg, oseg)
    assert abs(planes["normal"][0][2] + 1) < 1e-3 and abs(planes["d"][0] - 2000) < 2
    check_planes(planes, oplanes)


def test_batch_equals_single_and_oracle(drfe, orc):
    sel = [(0, 20260001), (1, 20260002), (2, 20260003), (0, 20260200)]
    data = [drfe.synth_frame(640, 480, s, seed, 1000.0) for s, seed in sel]
    depth = np.stack([d[1] for d in data])
    K = data[0][2]
    cpb = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0, max_batch=4)
    cpb.enqueue_depth(depth, *K)
    seg, planes, npl = cpb.download()
    cp1 = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0)
    o = orc.CapeOracle(480, 640, 20, 20, False, MC, 50.0)
    for f in range(4):
        r = cp1.process_depth(depth[f], *K)
        assert r[0] == npl[f] and np.array_equal(r[2], seg[f]) and same_planes(r[3], planes[f, :npl[f]])
        oseg, oplanes = o.process(o.depth_to_cloud(depth[f], *K))
        assert np.array_equal(seg[f], oseg)
        check_planes(planes[f, :npl[f]], oplanes)


def test_plane_detection_wrapper(drfe, orc):
    """PlaneDetection_CAPE::readDepthImage / runPlaneDetection (PlaneExtractor.cpp:101-191)."""
    _, depth, K = drfe.synth_frame(640, 480, 1, 20260090)
    pd = drfe.PlaneDetectionCAPE(PATCH_SIZE=20, MAX_MERGE_DIST=50.0)
    assert not pd.readDepthImage(depth.astype(np.float64), np.eye(3))     # wrong depth type -> false
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    assert pd.readDepthImage(depth, Km)
    pd.runPlaneDetection()
    o = orc.CapeOracle(480, 640, 20, 20, False, MC, 50.0)
    oseg, oplanes = o.process(o.depth_to_cloud(depth, *K))
    assert pd.nr_planes == len(oplanes) and pd.nr_cylinders == 0 and np.array_equal(pd.seg_output, oseg)


# ---------------------------------------------------------------- cylinder detection (CylinderSeg.cpp, config 5)
def check_cylinders(cp, o, depth, K, f=0, res=None):
    cloud = o.depth_to_cloud(depth, *K)
    oseg, oplanes, oncf, ocyls = o.process_full(cloud)
    pm, em = cp.grid_maps(f)
    opm, oem = o.grid_maps()
    assert np.array_equal(pm, opm) and np.array_equal(em, oem), "plane maps (planes re-fitted from extruded regions included)"
    cm, ce = cp.cyl_maps(f)
    ocm, oce = o.cyl_maps()
    assert np.array_equal(cm, ocm), "grid_cylinder_seg_map"
    assert np.array_equal(ce, oce), "eroded cylinder map (labels 50 + k)"
    return oseg, oplanes, oncf, ocyls


@pytest.mark.parametrize("w,h,scene,seed,unit", [
    (640, 480, 2, 20260100, 1000.0),
    (640, 480, 2, 20260105, 1000.0),
    (640, 480, 2, 20260204, 1000.0),     # a frame with six extruded sub-segments
    (640, 480, 2, 20260300, 1.0),        # metres: thresholds vacuous, many garbage cylinders
    (640, 480, 1, 20260502, 1000.0),
    (1280, 720, 2, 20260400, 1000.0),    # configs[4]: 1280x720, cylinders on
])
def test_cape_cylinders_parity(drfe, orc, w, h, scene, seed, unit):
    _, depth, K = drfe.synth_frame(w, h, scene, seed, unit)
    cp = drfe.CAPE(h, w, 20, 20, True, MC, 50.0)
    npl, ncyl, seg, planes, cyls = cp.process_depth(depth, *K)
    o = orc.CapeOracle(h, w, 20, 20, True, MC, 50.0)
    oseg, oplanes, oncf, ocyls = check_cylinders(cp, o, depth, K)
    assert npl == len(oplanes) and ncyl == oncf and len(cyls) == len(ocyls)
    assert np.array_equal(seg, oseg), "seg_output with cylinder labels"
    check_planes(planes, oplanes)
    if len(ocyls):
        # same operation sequence as the oracle: bit-identical (the bar would be 1e-5)
        for n in ("radius", "center", "axis"):
            assert np.array_equal(cyls[n], ocyls[n]), n


def test_cape_cylinders_golden(drfe):
    g = load_golden("cape_640x480_pillars_mm_cyl.npz")
    depth = g["depth_q"].astype(np.float32) * np.float32(1.0 / 5000.0) * g["unit"]
    cp = drfe.CAPE(480, 640, 20, 20, True, float(g["min_cos"]), float(g["max_merge"]))
    npl, ncyl, seg, planes, cyls = cp.process_depth(depth, *[float(v) for v in g["K"]])
    assert np.array_equal(seg, g["seg"]) and int(seg.max()) > 50
    assert ncyl == int(g["nr_cylinders_final"]) and len(cyls) == len(g["cyl_radius"])
    cm, ce = cp.cyl_maps()
    assert np.array_equal(cm, g["cyl_map"]) and np.array_equal(ce, g["cyl_eroded"])
    assert np.allclose(cyls["radius"], g["cyl_radius"], rtol=1e-6)
    assert np.allclose(cyls["center"], g["cyl_center"], rtol=1e-7, atol=1e-7)
    assert np.allclose(np.abs(cyls["axis"]), np.abs(g["cyl_axis"]), atol=1e-9)   # PCA sign is solver-defined
    assert np.allclose(planes["normal"], g["plane_normal"], atol=TOL, rtol=0) and np.allclose(planes["d"], g["plane_d"], rtol=1e-9)


def test_cape_cylinders_batch_matches_single(drfe, orc):
    """frames of a batch are independent: the rand() stream restarts per frame (App. B.9)"""
    seeds = [20260100, 20260105, 20260204, 20260101]
    frames = [drfe.synth_frame(640, 480, 2, s, 1000.0) for s in seeds]
    depth = np.stack([fr[1] for fr in frames])
    K = frames[0][2]
    cp = drfe.CAPE(480, 640, 20, 20, True, MC, 50.0, max_batch=4)
    cp.enqueue_depth(depth, *K)
    seg, planes, npl, ncyl, cyls, found = cp.download(with_cylinders=True)
    o = orc.CapeOracle(480, 640, 20, 20, True, MC, 50.0)
    for f in range(4):
        oseg, oplanes, oncf, ocyls = o.process_full(o.depth_to_cloud(depth[f], *K))
        assert np.array_equal(seg[f], oseg) and npl[f] == len(oplanes) and ncyl[f] == oncf and found[f] == len(ocyls)
        assert np.array_equal(cyls[f, :found[f]]["radius"], ocyls["radius"])
