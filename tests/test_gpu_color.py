"""The colour conversion of Tracking::GrabImageRGBD (reference src/Tracking.cc:194-207: cvtColor to gray before the Frame
is built) fused in front of the extractor.  CPU: the fixed-point model against cv2 on all 2^24 colours.
GPU: drfe_orb_enqueue_color's gray image bit-identical to cv2.cvtColor for RGB / BGR / RGBA / BGRA, and the extraction
that follows identical to extracting from that gray image."""
import numpy as np
import pytest

Q15 = (9798, 19235, 3735, 15)
Q14 = (4899, 9617, 1868, 14)


def model(rgb, q):
    r, g, b = [rgb[..., i].astype(np.int64) for i in range(3)]
    return ((r * q[0] + g * q[1] + b * q[2] + (1 << (q[3] - 1))) >> q[3]).astype(np.uint8)


def test_q15_model_is_cv2_on_every_colour():
    cv2 = pytest.importorskip("cv2")
    a = np.arange(256, dtype=np.uint8)
    full = np.stack(np.meshgrid(a, a, a, indexing="ij"), -1).reshape(4096, 4096, 3)
    assert np.array_equal(cv2.cvtColor(full, cv2.COLOR_RGB2GRAY), model(full, Q15))
    assert np.array_equal(cv2.cvtColor(full, cv2.COLOR_BGR2GRAY), model(full[..., ::-1], Q15))
    d = model(full, Q14).astype(int) - model(full, Q15).astype(int)
    assert np.abs(d).max() == 1 and 0 < (d != 0).mean() < 0.01          # the two OpenCV generations differ by one level on < 1 %


def colour_frames(drfe, n, channels, seed):
    """synthetic frames tinted per channel so that R, G and B differ"""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 480, 640, channels), np.uint8)
    for f in range(n):
        gray, _, _ = drfe.synth_frame(640, 480, f % 3, 20260600 + f)
        for c in range(3):
            out[f, ..., c] = np.clip(gray.astype(np.int32) * rng.uniform(0.6, 1.0) + rng.integers(-12, 13, gray.shape), 0, 255)
        if channels == 4:
            out[f, ..., 3] = rng.integers(0, 256, gray.shape)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("channels,rgb", [(3, True), (3, False), (4, True), (4, False)])
def test_gpu_colour_input(drfe, orc, channels, rgb):
    cv2 = pytest.importorskip("cv2")
    B = 2
    px = colour_frames(drfe, B, channels, 5 + channels)
    code = {(3, True): cv2.COLOR_RGB2GRAY, (3, False): cv2.COLOR_BGR2GRAY, (4, True): cv2.COLOR_RGBA2GRAY, (4, False): cv2.COLOR_BGRA2GRAY}[(channels, rgb)]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    ex.enqueue_color(px, rgb_order=rgb, coeffs=0)
    kps, desc, cnt = ex.download()
    for f in range(B):
        want = cv2.cvtColor(px[f], code)
        assert np.array_equal(ex.get_gray(f), want)
        rk, rd = orc.OrbOracle(1000).extract(want)
        assert cnt[f] == len(rk) and all(np.array_equal(kps[f, :cnt[f]][n], rk[n]) for n in rk.dtype.names)
        assert np.array_equal(desc[f, :cnt[f]], rd)
    # the OpenCV <= 3.4 constants
    ex.enqueue_color(px, rgb_order=rgb, coeffs=1)
    ex.sync()
    src = px[0][..., :3] if rgb else px[0][..., 2::-1]
    assert np.array_equal(ex.get_gray(0), model(src, Q14))


@pytest.mark.gpu
def test_gpu_colour_strided_rows(drfe):
    """a colour frame that is a view into a wider buffer (cv::Mat ROI): row stride larger than width * channels"""
    cv2 = pytest.importorskip("cv2")
    big = np.zeros((1, 480, 700, 3), np.uint8)
    big[:, :, :640] = colour_frames(drfe, 1, 3, 9)
    view = big[:, :, :640]
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    ex.enqueue_color(view, rgb_order=True)
    ex.sync()
    assert np.array_equal(ex.get_gray(0), cv2.cvtColor(np.ascontiguousarray(view[0]), cv2.COLOR_RGB2GRAY))
