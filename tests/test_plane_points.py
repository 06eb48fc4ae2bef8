"""Per-plane point lists (plane_cloud of PlaneDetection_CAPE::runPlaneDetection, reference
src/PlaneExtractor.cpp:165-190): the oracle's vectorised gather against a literal pixel loop on the CPU, and
drfe_cape_plane_points (device gather, C ABI) against the oracle on the GPU — bit-exact, order included."""
import numpy as np
import pytest

MC = float(np.float32(np.cos(np.pi / 12)))


def literal_plane_cloud(depth, K, seg, nr_planes):
    """the reference's loops: cloud_array in double -> float (:117-127), then code > 0 -> push_back (:176-190)"""
    fx, fy, cx, cy = (np.float32(v) for v in K)
    H, W = depth.shape
    out = [[] for _ in range(nr_planes)]
    for i in range(H):
        for j in range(W):
            code = int(seg[i, j])
            if code > 0:
                z = float(depth[i, j])
                x = (float(j) - float(cx)) * z / float(fx)
                y = (float(i) - float(cy)) * z / float(fy)
                out[code - 1].append((np.float32(x), np.float32(y), np.float32(z)))
    return [np.array(p, np.float32).reshape(-1, 3) for p in out]


def test_oracle_plane_points_literal(drfe, orc):
    _, depth, K = drfe.synth_frame(160, 120, 1, 20260005, 1000.0)
    o = orc.CapeOracle(120, 160, 10, 10, False, MC, 50.0)
    cloud = o.depth_to_cloud(depth, *K)
    seg, planes = o.process(cloud)
    assert len(planes) >= 1 and (seg > 0).sum() > 1000
    got = o.plane_points(cloud, seg, len(planes))
    want = literal_plane_cloud(depth, K, seg, len(planes))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert sum(len(a) for a in got) == int((seg > 0).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,scene,seed,unit,cell", [
    (640, 480, 0, 20260000, 1.0, 20),
    (640, 480, 2, 20260100, 1000.0, 20),
    (640, 480, 1, 20260012, 1000.0, 10),
    (1280, 720, 2, 20260140, 1000.0, 20),
])
def test_gpu_plane_points(drfe, orc, w, h, scene, seed, unit, cell):
    _, depth, K = drfe.synth_frame(w, h, scene, seed, unit)
    cp = drfe.CAPE(h, w, cell, cell, False, MC, 50.0)
    npl, _, seg, planes, _ = cp.process_depth(depth, *K)
    o = orc.CapeOracle(h, w, cell, cell, False, MC, 50.0)
    cloud = o.depth_to_cloud(depth, *K)
    oseg, oplanes = o.process(cloud)
    assert npl == len(oplanes) and np.array_equal(seg, oseg)
    want = o.plane_points(cloud, oseg, npl)
    pts, offs = cp.plane_points()
    assert offs[0, 0] == 0 and offs[0, npl] == int((oseg > 0).sum())
    for p in range(npl):
        got = pts[0, offs[0, p]:offs[0, p + 1]]
        assert got.shape == want[p].shape and np.array_equal(got, want[p]), "plane %d" % p


@pytest.mark.gpu
def test_gpu_plane_points_batch_and_cylinder_labels(drfe, orc):
    """a batch of frames (each frame's lists are independent), and cylinder labels (51+) are not gathered"""
    B, w, h = 5, 640, 480
    frames = [drfe.synth_frame(w, h, 2, 20260200 + 7 * i, 1000.0) for i in range(B)]
    depth = np.stack([f[1] for f in frames])
    K = frames[0][2]
    cp = drfe.CAPE(h, w, 20, 20, True, MC, 50.0, max_batch=B)
    cp.enqueue_depth(depth, *K, nframes=B)
    seg, planes, npl = cp.download()[:3]
    pts, offs = cp.plane_points(B)
    o = orc.CapeOracle(h, w, 20, 20, True, MC, 50.0)
    for f in range(B):
        cloud = o.depth_to_cloud(depth[f], *K)
        want = o.plane_points(cloud, seg[f], int(npl[f]))
        assert offs[f, npl[f]] == int(((seg[f] > 0) & (seg[f] <= npl[f])).sum())
        for p in range(int(npl[f])):
            assert np.array_equal(pts[f, offs[f, p]:offs[f, p + 1]], want[p]), (f, p)
