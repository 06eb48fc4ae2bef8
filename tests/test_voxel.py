"""The 5 cm pcl::VoxelGrid filter Frame::ComputePlanes_CAPE applies to every plane_cloud (reference src/Frame.cc:1121-1125)
and the 1/3-resolution cloud it builds for the surface normals (:1153-1172).  CPU: the C++ restatement of PCL 1.9's
VoxelGrid::applyFilter (oracle.voxel_grid) against an independent numpy formulation.  GPU: drfe_cape_plane_points_voxel and
drfe_cape_third_cloud against the oracle — bit for bit, order included."""
import numpy as np
import pytest

MC = float(np.float32(np.cos(np.pi / 12)))


def numpy_voxel_grid(pts, leaf):
    """independent formulation: leaf indices by numpy, leaves via np.unique, sums in ascending input order in float32"""
    f32 = np.float32
    inv = f32(1.0) / f32(leaf)
    mn, mx = pts.min(0), pts.max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    div = np.floor(mx * inv).astype(np.int64) - min_b + 1
    ijk = (np.floor(pts * inv) - min_b.astype(f32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    out = []
    for u in np.unique(idx):
        acc = np.zeros(3, f32)
        for p in pts[idx == u]:
            acc = (acc + p).astype(f32)
        out.append(acc / f32((idx == u).sum()))
    return np.array(out, f32).reshape(-1, 3)


def test_voxel_grid_restatement(orc):
    rng = np.random.default_rng(5)
    pts = (rng.random((6000, 3)) * np.array([1.5, 1.0, 0.08]) + np.array([-0.7, -0.5, 1.9])).astype(np.float32)
    pts[::97] = 0                                                 # pixels without depth sit at the origin
    got, unfiltered = orc.voxel_grid(pts, 0.05)
    want = numpy_voxel_grid(pts, 0.05)
    assert not unfiltered and got.shape == want.shape and np.array_equal(got, want)
    assert len(got) < len(pts) / 3
    one, _ = orc.voxel_grid(pts[:1], 0.05)
    assert np.array_equal(one, pts[:1])
    none, _ = orc.voxel_grid(np.zeros((0, 3), np.float32), 0.05)
    assert len(none) == 0
    # millimetre coordinates: the bounding box has more than INT32_MAX leaves of 0.05 -> PCL returns the input
    mm, unfiltered = orc.voxel_grid(pts * np.float32(1000), 0.05)
    assert unfiltered and np.array_equal(mm, pts * np.float32(1000))


def test_third_cloud_restatement(drfe, orc):
    _, depth, K = drfe.synth_frame(640, 480, 1, 20260021)
    got = orc.third_cloud(depth, *K, 3.0)
    fx, fy, cx, cy = (np.float32(v) for v in K)
    assert got.shape == (160, 214, 3)
    for (r, c) in [(0, 0), (10, 17), (159, 213), (80, 100)]:
        d = depth[3 * r, 3 * c]
        z = np.float32(0) if d > np.float32(3.0) else d
        want = ((np.float32(3 * c) - cx) * z / fx, (np.float32(3 * r) - cy) * z / fy, z)
        assert tuple(got[r, c]) == tuple(np.float32(v) for v in want)
    assert (got[..., 2] == 0).any() and (got[..., 2] > 0).any()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,scene,seed,unit,cell,leaf", [
    (640, 480, 0, 20260000, 1.0, 20, 0.05),
    (640, 480, 1, 20260012, 1.0, 20, 0.05),
    (640, 480, 2, 20260100, 1.0, 10, 0.02),
    (1280, 720, 2, 20260140, 1.0, 20, 0.05),
    (640, 480, 1, 20260012, 1000.0, 20, 50.0),                   # millimetres with a 50 mm leaf
    (640, 480, 1, 20260012, 1000.0, 20, 0.05),                   # millimetres with the reference's 0.05: unfiltered, like PCL
])
def test_gpu_voxel_filter(drfe, orc, w, h, scene, seed, unit, cell, leaf):
    _, depth, K = drfe.synth_frame(w, h, scene, seed, unit)
    cp = drfe.CAPE(h, w, cell, cell, False, MC, 50.0)
    npl, _, seg, planes, _ = cp.process_depth(depth, *K)
    o = orc.CapeOracle(h, w, cell, cell, False, MC, 50.0)
    cloud = o.depth_to_cloud(depth, *K)
    lists = o.plane_points(cloud, seg, npl)
    pts, offs = cp.plane_points_voxel(leaf)
    assert offs[0, 0] == 0
    total = 0
    for p in range(npl):
        want, _ = orc.voxel_grid(lists[p], leaf)
        got = pts[0, offs[0, p]:offs[0, p + 1]]
        assert got.shape == want.shape and np.array_equal(got, want), "plane %d" % p
        total += len(want)
    assert offs[0, npl] == total
    raw, roffs = cp.plane_points()                                 # the unfiltered lists are still what they were
    assert roffs[0, npl] == int((seg > 0).sum())


@pytest.mark.gpu
def test_gpu_voxel_filter_batch_and_third_cloud(drfe, orc):
    B, w, h = 6, 640, 480
    frames = [drfe.synth_frame(w, h, i % 3, 20260400 + 5 * i) for i in range(B)]
    depth = np.stack([f[1] for f in frames])
    K = frames[0][2]
    cp = drfe.CAPE(h, w, 20, 20, False, MC, 50.0, max_batch=B)
    cp.enqueue_depth(depth, *K, nframes=B)
    seg, planes, npl = cp.download()[:3]
    pts, offs = cp.plane_points_voxel(0.05, B)
    o = orc.CapeOracle(h, w, 20, 20, False, MC, 50.0)
    for f in range(B):
        lists = o.plane_points(o.depth_to_cloud(depth[f], *K), seg[f], int(npl[f]))
        for p in range(int(npl[f])):
            want, _ = orc.voxel_grid(lists[p], 0.05)
            assert np.array_equal(pts[f, offs[f, p]:offs[f, p + 1]], want), (f, p)
    third = cp.third_cloud(3.0, B)
    for f in range(B):
        assert third[f].tobytes() == orc.third_cloud(depth[f], *K, 3.0).tobytes()
    # raw 16-bit depth input gives the same 1/3 cloud
    q = np.rint(depth * 5000).astype(np.uint16)
    cp.enqueue_depth_u16(q, float(np.float32(1.0 / 5000.0)), *K)
    cp.download()
    assert cp.third_cloud(3.0, B).tobytes() == third.tobytes()
    with pytest.raises(drfe.DrfeError) as e:
        cp.plane_points_voxel(0.05, B, cap_per_frame=10)
    assert e.value.code == drfe.ERR_CAPACITY
