"""GPU: the capacity limits of the C ABI.  Either a limit cannot be reached by any input (the FAST candidate arena is
sized for the densest image strict 3x3 non-maximum suppression allows, and the test feeds exactly that image), or
reaching it returns DRFE_ERR_CAPACITY with a text — never a silent truncation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MC = float(np.float32(np.cos(np.pi / 12)))


def dense_corner_image(w, h, seed=3):
    """a 2 x 6 tile of dark pixels in a bright ground plus +-3 of noise: one kept FAST maximum per 4 pixels on level 0
    (found by search with the CPU oracle) — the most strict 3x3 non-maximum suppression can leave, twice what the
    kernel's shared-memory list holds"""
    rng = np.random.default_rng(seed)
    tile = np.array([[231, 231, 0, 231, 231, 231], [0, 231, 231, 231, 0, 231]], np.uint8)
    img = np.tile(tile, (h // 2 + 1, w // 6 + 1))[:h, :w].astype(np.int32)
    return np.clip(img + rng.integers(-3, 4, img.shape), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("size", [(640, 480), (752, 480)])
def test_fast_arena_holds_the_densest_possible_image(drfe, orc, size):
    w, h = size
    img = dense_corner_image(w, h)
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, w, h)
    kps, desc = ex(img, None)                                   # no DRFE_ERR_CAPACITY: the shared list overflows into the arena
    o = orc.OrbOracle(1000)
    rk, rd = o.extract(img)
    assert len(kps) == len(rk) > 500
    for f in ("x", "y", "response", "octave", "angle"):
        assert np.array_equal(kps[f], rk[f]), f
    ham = np.unpackbits(desc ^ rd, axis=1).sum(1)
    assert (ham == 0).mean() >= 0.995 and ham.max() <= 8
    for lvl in (0, 3, 7):
        got, ref = ex.candidates(0, lvl), o.candidates(lvl)
        assert len(got) == len(ref) and np.array_equal(np.unique(got, axis=0), np.unique(ref, axis=0))
    w0, h0 = ex.level_size(0)
    assert len(ex.candidates(0, 0)) > 0.2 * (w0 - 38) * (h0 - 38)   # far denser than the shared-memory list's one per 8 px


def test_download_caps_are_enforced(drfe):
    gray, depth, K = drfe.synth_frame(640, 480, 1, 20260777)
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=2)
    g2 = np.stack([gray, gray])
    ex.enqueue(g2)
    with pytest.raises(drfe.DrfeError) as e:
        ex.download(kps=np.empty((2, 500), drfe.KP_DTYPE), desc=np.empty((2, 500, 32), np.uint8))   # 1000+ keypoints, 500 slots
    assert e.value.code == drfe.ERR_CAPACITY
    with pytest.raises(drfe.DrfeError) as e:
        ex.extract_batch(g2, kps=np.empty((2, 500), drfe.KP_DTYPE), desc=np.empty((2, 500, 32), np.uint8))
        ex.finish_batch()
    assert e.value.code == drfe.ERR_CAPACITY
    kps, desc, cnt = ex.extract_batch(g2)                       # the handle is usable afterwards
    ex.finish_batch()
    assert cnt.min() >= 1000
    cp = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0)
    cp.enqueue_depth(depth[None], *K)
    with pytest.raises(drfe.DrfeError) as e:
        cp.download(planes=np.zeros((1, 1), drfe.PLANE_DTYPE))  # the room has more than one plane
    assert e.value.code == drfe.ERR_CAPACITY
    seg, planes, npl = cp.download()
    assert npl[0] > 1
    with pytest.raises(drfe.DrfeError) as e:
        cp.plane_points(cap_per_frame=100)                      # far fewer point slots than labelled pixels
    assert e.value.code == drfe.ERR_CAPACITY
    with pytest.raises(drfe.DrfeError) as e:
        cp.plane_points(plane_cap=1)
    assert e.value.code == drfe.ERR_CAPACITY
    pts, offs = cp.plane_points()
    assert offs[0, npl[0]] == (seg[0] > 0).sum()


def test_matchers_refuse_a_stale_frame_post(drfe):
    """the matchers work on what drfe_orb_frame_post left on the device; a new enqueue replaces the keypoints, so they
    must ask for a new frame_post instead of pairing old undistorted keys with new descriptors"""
    gray, depth, K = drfe.synth_frame(640, 480, 1, 20260778)
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    ex.enqueue(gray[None])
    p = ex.frame_params(*K, [0, 0, 0, 0, 0], 40.0)
    ex.frame_post(p, depth[None])
    q = np.zeros((1, 4), drfe.QUERY_DTYPE); q["x"] = 320; q["y"] = 240; q["r"] = 10; q["max_level"] = 7
    d = np.zeros((1, 4, 32), np.uint8)
    ex.search_by_projection(q, d)
    ex.enqueue(gray[None])
    with pytest.raises(drfe.DrfeError) as e:
        ex.search_by_projection(q, d)
    assert e.value.code == drfe.ERR_STATE
    ex.frame_post(p, depth[None])
    ex.search_by_projection(q, d)


def test_two_live_handles_of_different_geometry(drfe, orc):
    """the dynamic shared-memory limit of a kernel belongs to the device, not to a handle: creating a smaller handle must
    not shrink the limit the larger one launches with"""
    g_big, d_big, K = drfe.synth_frame(1280, 720, 2, 20260779, 1000.0)
    g_small, d_small, Ks = drfe.synth_frame(320, 240, 0, 20260780)
    big = drfe.ORBextractor(2000, 1.2, 8, 20, 7, 1280, 720)
    cbig = drfe.CAPE(720, 1280, 20, 20, False, MC, 50.0)
    k1, _ = big(g_big, None)
    small = drfe.ORBextractor(300, 1.2, 4, 20, 7, 320, 240)
    csmall = drfe.CAPE(240, 320, 10, 10, False, MC, 50.0)
    ks, _ = small(g_small, None)
    csmall.process_depth(d_small, *Ks)
    k2, _ = big(g_big, None)                                    # launches again after the smaller handles were created
    n2 = cbig.process_depth(d_big, *K)[0]
    assert k1.tobytes() == k2.tobytes() and len(ks) >= 300 and n2 >= 1
