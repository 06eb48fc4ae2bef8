"""CPU: pins the oracle's restated OpenCV primitives against the real OpenCV (cv2) and the
whole C++ oracle against the independent cv2-based Python restatement (oracle/py_ref.py).
Skipped where cv2 is not importable."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _rand_img(rng, h, w, smooth=True):
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if smooth:
        a = cv2.GaussianBlur(a, (5, 5), 1.2)
        a = cv2.normalize(a, None, 0, 255, cv2.NORM_MINMAX)
    return a


def test_resize_chain_bit_exact(orc):
    rng = np.random.default_rng(1)
    for (w, h) in [(640, 480), (1280, 720), (333, 251)]:
        img = _rand_img(rng, h, w)
        scale = np.float32(1.0)
        for _ in range(7):
            scale = np.float32(float(scale) * float(np.float32(1.2)))
            inv = np.float32(1.0) / scale
            dw, dh = int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))
            ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
            got = orc.resize_linear(img, dw, dh)
            assert np.array_equal(ref, got), (w, h, dw, dh)
            img = ref


def test_border_reflect101(orc):
    rng = np.random.default_rng(2)
    img = _rand_img(rng, 67, 91, smooth=False)
    assert np.array_equal(orc.border101(img, 19), cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("threshold", [20, 7])
def test_fast_nms_matches_cv2(orc, threshold):
    rng = np.random.default_rng(3 + threshold)
    det = cv2.FastFeatureDetector_create(threshold, True)
    for trial in range(40):
        h, w = int(rng.integers(7, 48)), int(rng.integers(7, 48))
        img = _rand_img(rng, h, w, smooth=bool(trial % 2))
        ref = np.array([(k.pt[0], k.pt[1], k.response) for k in det.detect(img)], np.float32).reshape(-1, 3)
        got = orc.fast9_nms(img, threshold)
        assert np.array_equal(ref, got), (trial, h, w)      # same keypoints, order and responses


def test_gaussian7_matches_cv2(orc):
    rng = np.random.default_rng(5)
    for (h, w) in [(480, 640), (134, 179), (9, 11)]:
        img = _rand_img(rng, h, w, smooth=False)
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(orc.gaussian7(img), ref)


def test_fast_atan2_matches_cv2(orc):
    rng = np.random.default_rng(6)
    for _ in range(5000):
        y, x = int(rng.integers(-2_000_000, 2_000_000)), int(rng.integers(-2_000_000, 2_000_000))
        assert orc.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x))), (y, x)


def test_orb_oracle_equals_cv2_restatement(orc, drfe):
    from oracle import py_ref
    gray, _, _ = drfe.synth_frame(416, 320, 1, 20260011)
    o = orc.OrbOracle(600)
    kps, desc = o.extract(gray)
    R = py_ref.OrbRef(600).extract(gray, keep_intermediates=True)
    for l in range(8):
        assert np.array_equal(o.level(l, True), R["bordered"][l])
        assert np.array_equal(o.candidates(l), np.array(R["cands"][l], np.float32).reshape(-1, 3))
    assert len(kps) == len(R["kps"])
    for f in ("x", "y", "size", "angle", "response", "octave"):
        assert np.array_equal(kps[f], R["kps"][f]), f
    assert np.array_equal(desc, R["desc"])


def test_cape_oracle_equals_numpy_restatement(orc, drfe):
    from oracle import py_ref
    _, depth, K = drfe.synth_frame(320, 240, 1, 20260044, 1000.0)
    mc = float(np.float32(np.cos(np.pi / 12)))
    o = orc.CapeOracle(240, 320, 20, 20, False, mc, 50.0)
    cloud = o.depth_to_cloud(depth, *K)
    assert np.array_equal(cloud, py_ref.depth_to_cloud(depth, *K, 20, 20))
    seg, planes = o.process(cloud)
    seg2, final2, grid2, pm2, em2 = py_ref.cape_process(cloud, 240, 320, 20, 20, mc, 50.0)
    pm, em = o.grid_maps()
    assert np.array_equal(pm, pm2) and np.array_equal(em, em2) and np.array_equal(seg, seg2)
    assert len(planes) == len(final2)
    for p, q in zip(planes, final2):
        assert np.allclose(p["normal"], q.normal, atol=1e-9) and abs(p["d"] - q.d) < 1e-9 * max(1, abs(q.d))
