"""cv::resize(im, IM, Size(640,480)) / cv::resize(depthmap, Depthmap, Size(640,480)) of System::TrackRGBD (reference
src/System.cc:325-329).  CPU: the numpy restatement (oracle.resize_input) is pinned against cv2 — bit for bit for 8U with
IPP on and off, and for 16U / 32F with IPP off (OpenCV's own arithmetic; IPP's resize kernel differs in the last bit and
is declared as not the model).  GPU: drfe_resize = the restatement, bit for bit."""
import numpy as np
import pytest

SIZES = [(848, 480), (1280, 720), (1280, 960), (960, 540), (641, 481), (640, 480), (320, 240), (500, 375), (640, 240)]


def make(kind, w, h, seed):
    rng = np.random.default_rng(seed)
    if kind == "rgb":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    if kind == "rgba":
        return rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if kind == "gray":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    d16 = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    d16[rng.random((h, w)) < 0.02] = 0                             # dropouts, like a sensor's depth map
    return d16 if kind == "u16" else (d16.astype(np.float32) * np.float32(1.0 / 5000.0)).astype(np.float32)


@pytest.mark.parametrize("size", SIZES)
def test_restatement_is_cv2_resize(orc, size):
    cv2 = pytest.importorskip("cv2")
    w, h = size
    had = cv2.ipp.useIPP()
    try:
        for ipp in (True, False):
            cv2.ipp.setUseIPP(ipp)
            for kind in ("rgb", "gray", "rgba"):
                src = make(kind, w, h, 7)
                if not (ipp and (w < 640 or h < 480)):               # 8U upscaling goes through IPP when it is on
                    assert np.array_equal(orc.resize_input(src, 640, 480), cv2.resize(src, (640, 480))), (kind, ipp)
        cv2.ipp.setUseIPP(False)
        for kind in ("u16", "f32"):
            src = make(kind, w, h, 8)
            assert np.array_equal(orc.resize_input(src, 640, 480), cv2.resize(src, (640, 480))), kind
    finally:
        cv2.ipp.setUseIPP(had)


@pytest.mark.gpu
@pytest.mark.parametrize("size", SIZES + [(1920, 1080)])
def test_device_resize_is_the_restatement(drfe, orc, size):
    w, h = size
    rs = drfe.Resizer(w, h, 640, 480, max_batch=3)
    for kind in ("rgb", "gray", "rgba", "u16", "f32"):
        batch = np.stack([make(kind, w, h, 20 + i) for i in range(3)])
        got = rs(batch)
        for i in range(3):
            assert got[i].tobytes() == orc.resize_input(batch[i], 640, 480).tobytes(), (kind, i)
    with pytest.raises(drfe.DrfeError):
        rs(np.zeros((4, h, w), np.uint8))                          # more frames than max_batch
    with pytest.raises(drfe.DrfeError):
        rs(np.zeros((1, h, w, 2), np.uint8))                       # 2 channels


@pytest.mark.gpu
def test_resized_input_feeds_the_extractors(drfe, orc):
    """the RealSense shape of DR-SLAM's input (848x480 colour + 16-bit depth) through resize -> cvtColor -> ORB and resize ->
    depth scaling -> CAPE equals the oracle run on the oracle's resized images"""
    w, h = 848, 480
    gray, depth, K = drfe.synth_frame(w, h, 1, 20260333)
    rgb = np.stack([gray, gray, gray], axis=2)
    q = np.rint(depth * 5000).astype(np.uint16)
    rs = drfe.Resizer(w, h, 640, 480)
    rgb2, q2 = rs(rgb[None])[0], rs(q[None])[0]
    assert np.array_equal(rgb2, orc.resize_input(rgb, 640, 480)) and np.array_equal(q2, orc.resize_input(q, 640, 480))
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    ex.enqueue_color(rgb2[None])
    kps, desc, cnt = ex.download()
    g2 = ex.get_gray(0)
    rk, rd = orc.OrbOracle(1000).extract(g2)
    assert cnt[0] == len(rk) and kps[0, :cnt[0]].tobytes() == rk.tobytes()
    mc = float(np.float32(np.cos(np.pi / 12)))
    cp = drfe.CAPE(480, 640, 20, 20, False, mc, 50.0)
    fac = np.float32(1.0 / 5000.0)
    sx, sy = 640.0 / w, 480.0 / h
    K2 = (K[0] * sx, K[1] * sy, K[2] * sx, K[3] * sy)
    cp.enqueue_depth_u16(q2[None], float(fac), *K2)
    seg, planes, npl = cp.download()
    o = orc.CapeOracle(480, 640, 20, 20, False, mc, 50.0)
    oseg, opl = o.process(o.depth_to_cloud(q2.astype(np.float32) * fac, *K2))
    assert npl[0] == len(opl) and np.array_equal(seg[0], oseg)
