"""CPU: the oracle's restatement of the per-frame steps after extraction —
UndistortKeyPoints, ComputeImageBounds, ComputeStereoFromRGBD, AssignFeaturesToGrid (Frame.cc:835-911,
224-237) — pinned against cv2.undistortPoints (the OpenCV routine the reference calls) and a literal
numpy restatement of the reference loops."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

TUM1 = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989,
            dist=[0.262383, -0.953104, -0.005358, 0.002628, 1.163314], bf=40.0)      # Examples/RGB-D/TUM1.yaml
ICL = dict(fx=481.2, fy=-480.0, cx=319.5, cy=239.5, dist=[0.0, 0.0, 0.0, 0.0, 0.0], bf=40.0)  # ICL.yaml: no distortion


def cv_undistort(pts, c):
    K = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]], np.float32)
    return cv2.undistortPoints(pts.reshape(-1, 1, 2), K, np.array(c["dist"], np.float32), None, K).reshape(-1, 2)


def test_undistort_point_is_cv_undistortPoints(orc):
    import ctypes as C
    p = orc.frame_params(TUM1["fx"], TUM1["fy"], TUM1["cx"], TUM1["cy"], TUM1["dist"], TUM1["bf"], 640, 480)
    pts = np.random.RandomState(7).uniform([0, 0], [640, 480], (3000, 2)).astype(np.float32)
    ref = cv_undistort(pts, TUM1)
    ou, ov = C.c_float(), C.c_float()
    got = np.zeros_like(pts)
    for i, (u, v) in enumerate(pts):
        orc.lib().orc_undistort_point(C.byref(p), float(u), float(v), C.byref(ou), C.byref(ov))
        got[i] = (ou.value, ov.value)
    assert np.array_equal(got, ref), "five fixed-point iterations in double, as cvUndistortPointsInternal"
    # ComputeImageBounds (Frame.cc:863-891)
    corners = cv_undistort(np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32), TUM1)
    assert p.min_x == min(corners[0, 0], corners[2, 0]) and p.max_x == max(corners[1, 0], corners[3, 0])
    assert p.min_y == min(corners[0, 1], corners[1, 1]) and p.max_y == max(corners[2, 1], corners[3, 1])


@pytest.mark.parametrize("calib", [TUM1, dict(TUM1, dist=[0.0] * 5)])
def test_frame_post_matches_reference_loops(orc, drfe, calib):
    gray, depth, _ = drfe.synth_frame(640, 480, 1, 20260042)   # host-only generator
    keys, _ = orc.OrbOracle(1000).extract(gray)
    p = orc.frame_params(calib["fx"], calib["fy"], calib["cx"], calib["cy"], calib["dist"], calib["bf"], 640, 480)
    ku, ur, kd, gc, gi = orc.frame_post(p, keys, depth)
    # literal restatement
    xy = np.stack([keys["x"], keys["y"]], 1)
    un = cv_undistort(xy, calib) if calib["dist"][0] != 0 else xy
    assert np.array_equal(np.stack([ku["x"], ku["y"]], 1), un)
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(ku[f], keys[f])
    d = depth[keys["y"].astype(np.int32), keys["x"].astype(np.int32)]
    assert np.array_equal(kd, np.where(d > 0, d, np.float32(-1)))
    with np.errstate(divide="ignore"):
        exp_ur = np.where(d > 0, un[:, 0] - np.float32(calib["bf"]) / d, np.float32(-1)).astype(np.float32)
    assert np.array_equal(ur, exp_ur)
    inv_w = np.float32(64) / np.float32(p.max_x - p.min_x)
    inv_h = np.float32(48) / np.float32(p.max_y - p.min_y)
    grid = [[[] for _ in range(48)] for _ in range(64)]
    for i in range(len(keys)):
        px = int(np.floor(abs(float((un[i, 0] - np.float32(p.min_x)) * inv_w)) + 0.5) * np.sign(float((un[i, 0] - np.float32(p.min_x)) * inv_w)))
        py = int(np.floor(abs(float((un[i, 1] - np.float32(p.min_y)) * inv_h)) + 0.5) * np.sign(float((un[i, 1] - np.float32(p.min_y)) * inv_h)))
        if 0 <= px < 64 and 0 <= py < 48:
            grid[px][py].append(i)
    assert np.array_equal(gc.reshape(64, 48), np.array([[len(c) for c in col] for col in grid], np.uint16))
    assert gi.tolist() == [i for col in grid for c in col for i in c]
    assert gc.sum() == len(gi) > 900


@pytest.mark.parametrize("calib,size", [(TUM1, (640, 480)), (ICL, (640, 480)), (TUM1, (1280, 720)),
                                        (dict(TUM1, dist=[-0.3, 0.12, 0.001, -0.002, 0.0]), (752, 480))])
def test_product_image_bounds_equal_oracle_and_cv2(drfe, orc, calib, size):
    """drfe_frame_image_bounds (Frame::ComputeImageBounds, Frame.cc:863-891) is host code of the product library and runs
    without a GPU: it must equal the oracle's and the four corners undistorted by cv2.undistortPoints"""
    import ctypes as C
    w, h = size
    p = drfe.FrameParams(calib["fx"], calib["fy"], calib["cx"], calib["cy"], (C.c_float * 5)(*calib["dist"]), calib["bf"], 0, 0, 0, 0)
    assert drfe.lib().drfe_frame_image_bounds(C.byref(p), w, h) == 0
    o = orc.frame_params(calib["fx"], calib["fy"], calib["cx"], calib["cy"], calib["dist"], calib["bf"], w, h)
    assert (p.min_x, p.max_x, p.min_y, p.max_y) == (o.min_x, o.max_x, o.min_y, o.max_y)
    if calib["dist"][0] != 0.0:
        c = cv_undistort(np.array([[0, 0], [w, 0], [0, h], [w, h]], np.float32), calib)
        assert p.min_x == min(c[0, 0], c[2, 0]) and p.max_x == max(c[1, 0], c[3, 0])
        assert p.min_y == min(c[0, 1], c[1, 1]) and p.max_y == max(c[2, 1], c[3, 1])
    else:
        assert (p.min_x, p.max_x, p.min_y, p.max_y) == (0.0, float(w), 0.0, float(h))
    assert drfe.lib().drfe_frame_image_bounds(None, w, h) == drfe.ERR_ARG
