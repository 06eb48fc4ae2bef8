"""GPU: the CUDA ORB path, called through the C ABI, against the CPU oracle and the golden
fixtures.  Bars (BASELINE.json north_star): pyramid bytes, FAST candidate sets (x, y, level,
response) and the quadtree-retained keypoints bit-exact; angles within 1e-3 rad (here: equal);
descriptors bit-identical on >= 99.5 % of keypoints, the rest within Hamming distance 8."""
import numpy as np
import pytest

from conftest import KP_FIELDS, ROOT, load_golden, sort_rows

pytestmark = pytest.mark.gpu

MAX_HAMMING = 8          # stated bound for the < 0.5 % of descriptors allowed to differ
ANGLE_TOL_DEG = 1e-3 * 180 / np.pi


def check_frame_against_oracle(ex, orc_o, gray, f, stages=True):
    orc_o.run(gray)
    if stages:
        for l in range(ex.nlevels):
            assert np.array_equal(ex.pyramid(f, l, bordered=True), orc_o.level(l, bordered=True)), "pyramid level %d" % l
            assert np.array_equal(sort_rows(ex.candidates(f, l)), sort_rows(orc_o.candidates(l))), "FAST level %d" % l
            ka, kb = ex.level_keypoints(f, l), orc_o.level_keypoints(l)
            assert len(ka) == len(kb), "quadtree count level %d" % l
            for n in ("x", "y", "response", "octave", "size"):
                assert np.array_equal(ka[n], kb[n]), "quadtree %s level %d" % (n, l)
            assert np.all(np.abs(ka["angle"] - kb["angle"]) <= ANGLE_TOL_DEG)
            if len(kb):
                assert np.array_equal(ex.blurred(f, l), orc_o.blurred(l)), "blur level %d" % l
    return orc_o.result()


def check_result(kps, desc, rk, rd):
    assert len(kps) == len(rk)
    for n in KP_FIELDS:
        if n == "angle":
            assert np.all(np.abs(kps[n] - rk[n]) <= ANGLE_TOL_DEG)
        else:
            assert np.array_equal(kps[n], rk[n]), n
    if len(rk):
        ham = np.unpackbits(desc ^ rd, axis=1).sum(1)
        assert (ham == 0).mean() >= 0.995, "identical descriptors: %.3f%%" % ((ham == 0).mean() * 100)
        assert ham.max() <= MAX_HAMMING


@pytest.mark.parametrize("w,h,nfeat,scene,seed", [
    (640, 480, 1000, 0, 20260000),      # BASELINE configs[0]/[1]: corridor frame
    (640, 480, 1000, 1, 20260077),
    (640, 480, 800, 2, 20260031),       # Realsense.yaml uses 800 features
    (320, 240, 500, 0, 20260005),
    (1280, 720, 2000, 2, 20260140),     # configs[4] geometry (2 quadtree roots)
    (752, 480, 1200, 1, 20260009),      # non-4:3 aspect, odd level sizes; level 0 is cut into two FAST segments
    (1920, 1080, 3000, 1, 20260201),    # wide levels: up to four 16-cell FAST segments per cell row
])
def test_orb_stage_parity(drfe, orc, w, h, nfeat, scene, seed):
    gray, _, _ = drfe.synth_frame(w, h, scene, seed)
    ex = drfe.ORBextractor(nfeat, 1.2, 8, 20, 7, w, h)
    kps, desc = ex(gray, None)
    assert ex.features_per_level() == orc.OrbOracle(nfeat).features_per_level()
    rk, rd = check_frame_against_oracle(ex, orc.OrbOracle(nfeat), gray, 0)
    check_result(kps, desc, rk, rd)
    ex.close()


@pytest.mark.parametrize("name", ["orb_320x240_corridor.npz", "orb_640x480_room.npz"])
def test_orb_matches_golden(drfe, name):
    g = load_golden(name)
    gray = g["gray"]
    ex = drfe.ORBextractor(int(g["nfeatures"]), 1.2, 8, 20, 7, gray.shape[1], gray.shape[0])
    kps, desc = ex(gray)
    assert np.array_equal(ex.pyramid(0, 3, bordered=True), g["pyr_level3"])
    for l in range(8):
        assert np.array_equal(sort_rows(ex.candidates(0, l)), sort_rows(g["cands_%d" % l].astype(np.float32)))
        lk, gl = ex.level_keypoints(0, l), g["lkp_%d" % l]
        assert len(lk) == len(gl)
        if len(gl):
            assert np.array_equal(np.stack([lk["x"], lk["y"], lk["response"]], 1), gl[:, :3])
    check_result(kps, desc, g["kps"], g["desc"])


def test_getters_mirror_reference(drfe, orc):
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7)
    assert ex.GetLevels() == 8 and abs(ex.GetScaleFactor() - 1.2) < 1e-6
    assert np.array_equal(ex.GetScaleFactors(), np.array(orc.OrbOracle(1000).scale_factors(), np.float32))
    assert np.allclose(ex.GetInverseScaleFactors() * ex.GetScaleFactors(), 1, atol=1e-6)
    assert np.array_equal(ex.GetScaleSigmaSquares(), ex.GetScaleFactors() ** 2)
    assert ex.features_per_level() == [217, 181, 151, 126, 105, 87, 73, 60]


def test_empty_and_wrong_inputs(drfe):
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7)
    assert ex(np.empty((0, 0), np.uint8)) == (None, None)        # reference: silent return (ORBextractor.cc:1046)
    with pytest.raises(AssertionError):
        ex(np.zeros((480, 640), np.float32))                     # reference asserts CV_8UC1 (:1050)
    with pytest.raises(drfe.DrfeError) as e:
        ex(np.zeros((240, 320), np.uint8))                       # size differs from the handle
    assert e.value.code == drfe.ERR_ARG


def test_flat_image_gives_no_keypoints(drfe, orc):
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7)
    kps, desc = ex(np.full((480, 640), 127, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    assert len(orc.OrbOracle(1000).extract(np.full((480, 640), 127, np.uint8))[0]) == 0


def test_low_contrast_triggers_min_threshold_fallback(drfe, orc):
    """Cells with no FAST(20) corner fall back to FAST(7) (ORBextractor.cc:812-816)."""
    gray, _, _ = drfe.synth_frame(640, 480, 0, 20260050)
    low = (96 + (gray.astype(np.int32) - 128) // 6).astype(np.uint8)      # contrast / 6
    low[:, :320] = gray[:, :320]                                          # half the cells keep contrast
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7)
    kps, desc = ex(low)
    o = orc.OrbOracle(1000)
    rk, rd = check_frame_against_oracle(ex, o, low, 0)
    check_result(kps, desc, rk, rd)
    c0 = o.candidates(0)
    assert (c0[:, 2] < 20).any() and (c0[:, 2] >= 20).any()               # both thresholds in play


@pytest.mark.parametrize("kind", ["noise", "dots", "steps"])
def test_dense_and_adversarial_corner_patterns(drfe, orc, kind):
    """Patterns that stress the FAST stages: uniform noise (corners of both polarities everywhere, equal-score
    neighbours, maxima on both sides of cell edges), isolated single-pixel dots on a ramp (many equal scores, each
    cell's NMS decided by strictness), and intensity steps near the thresholds (|d| == t must not count)."""
    rng = np.random.default_rng(11)
    if kind == "noise":
        img = rng.integers(0, 256, (480, 640)).astype(np.uint8)
        img[:, 320:] = (img[:, 320:] // 8 + 100).astype(np.uint8)          # right half: amplitude 32, thresholds matter
    elif kind == "dots":
        img = np.tile((np.arange(640) // 5).astype(np.uint8), (480, 1))
        ys, xs = rng.integers(20, 460, 4000), rng.integers(20, 620, 4000)
        img[ys, xs] = np.where(rng.random(4000) < 0.5, 255, 0)
    else:
        img = np.full((480, 640), 100, np.uint8)
        for k, d in enumerate((6, 7, 8, 19, 20, 21, 22, 40)):              # steps of exactly t-1, t, t+1 (t = 7, 20)
            img[60 * k:60 * k + 30, :] = 100 + d
            img[60 * k:60 * k + 30, ::17] = 100
            img[60 * k + 5:60 * k + 30:7, 3::13] = 100 - d
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    rk, rd = check_frame_against_oracle(ex, orc.OrbOracle(1000), img, 0)
    check_result(kps, desc, rk, rd)
    assert len(kps) > 0


def test_few_corners_sparse_quadtree(drfe, orc):
    """Far fewer candidates than nfeatures: every node ends with one key, nothing to drop."""
    img = np.full((480, 640), 100, np.uint8)
    rng = np.random.default_rng(7)
    for _ in range(40):
        y, x = int(rng.integers(40, 440)), int(rng.integers(40, 600))
        img[y:y + 9, x:x + 9] = 200
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    rk, rd = check_frame_against_oracle(ex, orc.OrbOracle(1000), img, 0)
    check_result(kps, desc, rk, rd)
    assert 0 < len(kps) < 1000


def test_strided_input_and_batch_equals_single(drfe, orc):
    frames = [drfe.synth_frame(640, 480, s, seed)[0] for s, seed in [(0, 20260001), (1, 20260002), (2, 20260003)]]
    big = np.zeros((3, 500, 704), np.uint8)
    big[:, :480, :640] = np.stack(frames)
    view = big[:, :480, :640]                                   # row stride 704, frame stride 500*704
    exb = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=3)
    exb.enqueue(view)
    kps, desc, counts = exb.download()
    ex1 = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    for f in range(3):
        k1, d1 = ex1(frames[f])
        assert counts[f] == len(k1)
        assert kps[f, :counts[f]].tobytes() == k1.tobytes() and np.array_equal(desc[f, :counts[f]], d1)
    rk, rd = orc.OrbOracle(1000).extract(frames[2])
    check_result(kps[2, :counts[2]], desc[2, :counts[2]], rk, rd)


def test_repeated_calls_are_idempotent(drfe):
    gray, _, _ = drfe.synth_frame(640, 480, 1, 20260021)
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7)
    a = ex(gray)
    ex(drfe.synth_frame(640, 480, 0, 20260022)[0])              # different frame in between
    b = ex(gray)
    assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])


def test_other_pyramid_settings(drfe, orc):
    gray, _, _ = drfe.synth_frame(640, 480, 1, 20260060)
    for (nf, sf, nl, ini, mn) in [(500, 1.5, 4, 25, 10), (1500, 1.1, 10, 15, 5), (300, 2.0, 3, 20, 7)]:
        ex = drfe.ORBextractor(nf, sf, nl, ini, mn)
        kps, desc = ex(gray)
        o = orc.OrbOracle(nf, sf, nl, ini, mn)
        rk, rd = check_frame_against_oracle(ex, o, gray, 0)
        check_result(kps, desc, rk, rd)
        ex.close()


def test_generic_pyramid_kernel_still_matches(drfe):
    """k_pyr_resize (the tile kernel kept as fallback for scale factors the streaming kernel cannot take) is forced
    with DRFE_PYR_GENERIC=1 in a child process and must produce the golden pyramid bytes and keypoints too"""
    import os
    import subprocess
    import sys
    code = (
        "import sys, zlib, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import drfe\n"
        "g = np.load(%r)\n"
        "ex = drfe.ORBextractor(int(g['nfeatures']), 1.2, 8, 20, 7, g['gray'].shape[1], g['gray'].shape[0])\n"
        "kps, desc = ex(g['gray'])\n"
        "assert all(np.uint32(zlib.crc32(ex.pyramid(0, l, bordered=True).tobytes())) == g['pyr_crc_%%d' %% l] for l in range(8))\n"
        "assert kps.tobytes() == g['kps'].tobytes() and np.array_equal(desc, g['desc'])\n"
        "print('generic ok')\n"
    ) % (os.path.join(ROOT, "dr-slam_b200"), ROOT, os.path.join(ROOT, "tests", "golden", "orb_640x480_room.npz"))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DRFE_PYR_GENERIC="1"), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "generic ok" in out.stdout, out.stderr[-2000:]
