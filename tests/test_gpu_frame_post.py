"""GPU: the per-frame steps after extraction through the C ABI against the CPU oracle:
mvKeysUn, mvuRight, mvDepth and the 64 x 48 feature grid must be bit-identical, with and without lens
distortion, for a single frame and for every frame of a batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TUM1 = (517.306408, 516.469215, 318.643040, 255.313989, [0.262383, -0.953104, -0.005358, 0.002628, 1.163314], 40.0)
NODIST = (525.0, 525.0, 319.5, 239.5, [0.0] * 5, 40.0)


def check_frame(orc, p_o, keys, depth, ku, ur, kd, gc, gi):
    oku, our, okd, ogc, ogi = orc.frame_post(p_o, keys, depth)
    n = len(keys)
    assert ku[:n].tobytes() == oku.tobytes(), "mvKeysUn"
    assert np.array_equal(ur[:n], our) and np.array_equal(kd[:n], okd), "mvuRight / mvDepth"
    assert np.array_equal(gc.reshape(-1), ogc), "mGrid sizes"
    assert np.array_equal(gi[:len(ogi)], ogi), "mGrid contents in push_back order"


@pytest.mark.parametrize("calib", [TUM1, NODIST])
def test_frame_post_single(drfe, orc, calib):
    gray, depth, _ = drfe.synth_frame(640, 480, 2, 20260314)
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480)
    ex.enqueue(gray[None])
    kps, _, cnt = ex.download()
    p = ex.frame_params(*calib)
    p_o = orc.frame_params(*calib, 640, 480)
    assert (p.min_x, p.max_x, p.min_y, p.max_y) == (p_o.min_x, p_o.max_x, p_o.min_y, p_o.max_y), "ComputeImageBounds"
    ku, ur, kd, gc, gi = ex.frame_post(p, depth[None])
    check_frame(orc, p_o, kps[0, :cnt[0]], depth, ku[0], ur[0], kd[0], gc[0], gi[0])
    assert (kd[0, :cnt[0]] > 0).mean() > 0.5 and gc[0].sum() > 900


def test_frame_post_batch_device_depth(drfe, orc):
    torch = pytest.importorskip("torch")
    B = 24
    data = [drfe.synth_frame(640, 480, i % 3, 20260700 + i) for i in range(B)]
    gray = np.stack([d[0] for d in data]); depth = np.stack([d[1] for d in data])
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    ex.enqueue(gray)
    kps, _, cnt = ex.download()
    p, p_o = ex.frame_params(*TUM1), orc.frame_params(*TUM1, 640, 480)
    d_depth = torch.from_numpy(depth).cuda()
    ku, ur, kd, gc, gi = ex.frame_post(p, d_depth.data_ptr(), mem_kind=drfe.MEM_DEVICE, row_stride=640, frame_stride=640 * 480)
    for f in range(B):
        check_frame(orc, p_o, kps[f, :cnt[f]], depth[f], ku[f], ur[f], kd[f], gc[f], gi[f])


def test_frame_post_shares_the_depth_of_the_cape_handle(drfe, orc):
    """Frame::Frame gives one imDepth to the plane extractor and to ComputeStereoFromRGBD: the depth a CAPE handle uploaded
    (float metres, or raw u16 scaled on the device like convertTo, Frame.cc:113-115) serves drfe_orb_frame_post too"""
    B = 6
    data = [drfe.synth_frame(640, 480, i % 3, 20260720 + i) for i in range(B)]
    gray = np.stack([d[0] for d in data]); depth = np.stack([d[1] for d in data])
    K = data[0][2]
    mc = float(np.float32(np.cos(np.pi / 12)))
    ex = drfe.ORBextractor(1000, 1.2, 8, 20, 7, 640, 480, max_batch=B)
    cp = drfe.CAPE(480, 640, 20, 20, False, mc, 50.0, max_batch=B)
    p = ex.frame_params(*TUM1)
    ex.enqueue(gray)
    cp.enqueue_depth(depth, *K)                                  # no host synchronisation in between: the streams order it
    got = ex.frame_post_shared_depth(p, cp)
    want = ex.frame_post(p, depth)
    for g, w in zip(got, want):
        assert g.tobytes() == w.tobytes()
    # raw 16-bit depth
    q = np.rint(depth * 5000).astype(np.uint16)
    fac = np.float32(1.0 / 5000.0)
    assert np.array_equal(q.astype(np.float32) * fac, depth)
    cp.enqueue_depth_u16(q, float(fac), *K)
    got = ex.frame_post_shared_depth(p, cp)
    for g, w in zip(got, want):
        assert g.tobytes() == w.tobytes()
    # after a pipelined batch call
    seg, planes, npl, _, _ = cp.process_depth_batch(q, *K, depth_factor=float(fac))
    cp.finish_batch()
    got = ex.frame_post_shared_depth(p, cp)
    for g, w in zip(got, want):
        assert g.tobytes() == w.tobytes()
    # a handle that was fed a cloud has no depth to share
    cloud = orc.CapeOracle(480, 640, 20, 20, False, mc, 50.0).depth_to_cloud(depth[0], *K)
    cp1 = drfe.CAPE(480, 640, 20, 20, False, mc, 50.0, max_batch=B)
    cp1.enqueue_cloud(np.stack([cloud] * B))
    with pytest.raises(drfe.DrfeError):
        ex.frame_post_shared_depth(p, cp1)
