"""The surface normals DR-SLAM takes from pcl::IntegralImageNormalEstimation on the 1/3-resolution cloud (reference src/Frame.cc:1174-1216).
PCL is not vendored in the reference and not installed here: oracle/normals_oracle.cpp restates PCL 1.9's published algorithm (parity
unpinned by PCL).  CPU: properties of the restatement — exact normals on planes seen in perspective, PCL's NaN pattern, the window rule,
depth discontinuities.  GPU: drfe_cape_third_cloud_normals is bit-identical to the restatement, NaN for NaN."""
import numpy as np
import pytest

MC = float(np.float32(np.cos(np.pi / 12)))


def plane_cloud(w, h, n, d, fx, fy, cx, cy):
    jj, ii = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    z = d / (n[0] * (jj - cx) / fx + n[1] * (ii - cy) / fy + n[2])
    return np.stack([(jj - cx) / fx * z, (ii - cy) / fy * z, z], -1).astype(np.float32)


def test_normals_of_planes_and_the_nan_border(orc):
    w, h = 214, 160
    for n, d in (((0.3, -0.2, 1.0), 2.0), ((0.0, 0.0, 1.0), 1.5), ((-0.5, 0.1, 0.8), 3.0)):
        cloud = plane_cloud(w, h, n, d, 175.0, 175.0, 106.5, 79.5)
        nr, dm = orc.integral_normals(cloud, with_distance_map=True)
        want = -np.array(n) / np.linalg.norm(n)                     # flipped towards the camera at the origin
        assert np.isnan(nr[:10]).all() and np.isnan(nr[-10:]).all() and np.isnan(nr[:, :10]).all() and np.isnan(nr[:, -10:]).all()
        assert not np.isnan(nr[10:-10, 10:-10]).any() and np.abs(nr[10:-10, 10:-10] - want).max() < 2e-6
        assert (dm == w + h).all()                                  # no depth discontinuity anywhere


def test_depth_discontinuity_and_window_rule(orc):
    w, h = 214, 160
    cloud = plane_cloud(w, h, (0.0, 0.0, 1.0), 2.0, 175.0, 175.0, 106.5, 79.5)
    near = plane_cloud(w, h, (0.2, 0.0, 1.0), 1.0, 175.0, 175.0, 106.5, 79.5)
    cloud[40:120, 60:150] = near[40:120, 60:150]                    # a box in front of the wall: jumps of ~1 m along its outline
    nr, dm = orc.integral_normals(cloud, with_distance_map=True)
    assert dm[40, 100] == 0 and dm[39, 100] == 0 and dm[80, 60] == 0 and dm[80, 59] == 0       # both sides of the jump are marked
    assert dm[80, 100] > 10 and dm[20, 30] > 10
    # chamfer distances: 1.0 per axial step, 1.4f per diagonal step, float additions in scan order
    assert dm[41, 100] == np.float32(1.0) and dm[42, 100] == np.float32(2.0) and dm[80, 63] == np.float32(3.0)
    inside = nr[60:100, 80:130]
    assert np.abs(inside - (-np.array([0.2, 0, 1.0]) / np.linalg.norm([0.2, 0, 1.0]))).max() < 2e-6
    # smoothing = min(distance, 10) > 2 is needed: next to the jump (distance <= 2) the normal is NaN, three steps away it is not
    assert np.isnan(nr[41, 100]).all() and np.isnan(nr[42, 100]).all() and not np.isnan(nr[43, 100]).any()
    # a zero-depth region (the far cull writes x = y = z = 0) is finite: PCL estimates there too; |n|^2 == 0 gives NaN
    cloud2 = cloud.copy(); cloud2[130:150, 20:200] = 0
    nr2 = orc.integral_normals(cloud2)
    assert np.isnan(nr2[138:142, 60:160]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("scene,seed,u16,clean", [(1, 20260042, False, True), (2, 20260100, False, False), (0, 20260007, True, True)])
def test_gpu_third_cloud_normals_equal_the_restatement(drfe, orc, scene, seed, u16, clean):
    from test_peac import clean_depth
    frames = [drfe.synth_frame(640, 480, scene, seed + i) for i in range(3)]
    K = frames[0][2]
    # the synthetic sensor drops 2 % of the pixels independently: every dropout is a depth jump, and with them almost every window is
    # rejected (the NaN pattern is then most of the result); the cleaned variant keeps coherent holes only, like a real sensor
    depth = np.stack([clean_depth(f[1], ((100, 140, 300, 420),)) if clean else f[1] for f in frames])
    cp = drfe.CAPE(480, 640, 20, 20, False, MC, 50.0, max_batch=3)
    if u16:
        q = np.rint(depth * 5000).astype(np.uint16)
        fac = float(np.float32(1.0 / 5000.0))
        cp.enqueue_depth_u16(q, fac, *K)
        depth = q.astype(np.float32) * np.float32(fac)
    else:
        cp.enqueue_depth(depth, *K)
    cp.download()
    dist_max = 10.0 if clean else 3.0                              # mMax_point_dist: with 3 m most of these rooms is culled to (0, 0, 0)
    cloud, normals = cp.third_cloud_normals(dist_max)
    for f in range(3):
        want_cloud = orc.third_cloud(depth[f], *K, dist_max)
        assert cloud[f].tobytes() == want_cloud.tobytes()
        want = orc.integral_normals(want_cloud, 0.05, 10.0)
        assert np.array_equal(np.isnan(normals[f]), np.isnan(want)), f
        ok = ~np.isnan(want)
        assert ok.mean() > (0.5 if clean else 0.01) and normals[f][ok].tobytes() == want[ok].tobytes(), f


@pytest.mark.gpu
def test_gpu_peac_handle_gives_the_same_cloud_and_normals(drfe, orc):
    """Frame::ComputePlanes (the PEAC path, Frame.cc:1044-1100) builds the same 1/3 cloud from imDepth = float(raw) * factor"""
    from test_peac import FAC, frame
    q, K = frame(drfe, 1, 20260012, ((100, 140, 300, 420),))
    pe = drfe.PEAC(640, 480)
    pe.enqueue(q[None], FAC, *K)
    pe.download()
    cloud, normals = pe.third_cloud_normals(10.0)
    depth = q.astype(np.float32) * np.float32(FAC)
    want_cloud = orc.third_cloud(depth, *K, 10.0)
    assert cloud[0].tobytes() == want_cloud.tobytes()
    want = orc.integral_normals(want_cloud)
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(normals[0]), ~ok) and ok.mean() > 0.5 and normals[0][ok].tobytes() == want[ok].tobytes()


def brute_force_normals(cloud, factor=0.05, smoothing=10.0):
    """an independent formulation of the same published algorithm: scatter-form change map, the raster passes as plain Python loops over a
    flat list (so the one-past-the-row reads are literal), window sums by numpy's own 2-D cumulative sums in float64"""
    h, w, _ = cloud.shape
    z = cloud[:, :, 2].astype(np.float32)
    change = np.full(h * w, 255, np.uint8)
    th = (np.float32(factor) * (np.abs(z) + np.float32(1.0)) * np.float32(2.0)).astype(np.float32)
    for r in range(h - 1):
        for c in range(w - 1):
            i = r * w + c
            if abs(float(z[r, c] - z[r, c + 1])) > th[r, c] or not np.isfinite(z[r, c]) or not np.isfinite(z[r, c + 1]):
                change[i] = change[i + 1] = 0
            if abs(float(z[r, c] - z[r + 1, c])) > th[r, c] or not np.isfinite(z[r, c]) or not np.isfinite(z[r + 1, c]):
                change[i] = change[i + w] = 0
    f32 = np.float32
    d = [f32(0.0) if v == 0 else f32(w + h) for v in change]
    one, diag = f32(1.0), f32(1.4)
    for r in range(1, h):
        pr, cu = (r - 1) * w, r * w
        for c in range(1, w):
            m = min(min(d[pr + c - 1] + diag, d[pr + c] + one), min(d[cu + c - 1] + one, d[pr + c + 1] + diag))
            if m < d[cu + c]:
                d[cu + c] = m
    for r in range(h - 2, -1, -1):
        nx, cu = (r + 1) * w, r * w
        for c in range(w - 2, -1, -1):
            m = min(min(d[nx + c - 1] + diag, d[nx + c] + one), min(d[cu + c + 1] + one, d[nx + c + 1] + diag))
            if m < d[cu + c]:
                d[cu + c] = m
    dm = np.array(d, np.float32).reshape(h, w)
    P = cloud.astype(np.float32)
    dx = np.zeros((h, w, 3), np.float32); dy = np.zeros((h, w, 3), np.float32)
    dx[1:-1, 1:-1] = P[1:-1, 2:] - P[1:-1, :-2]
    dy[1:-1, 1:-1] = P[2:, 1:-1] - P[:-2, 1:-1]
    Ix = np.zeros((h + 1, w + 1, 3)); Iy = np.zeros((h + 1, w + 1, 3))
    Ix[1:, 1:] = dx.astype(np.float64).cumsum(0).cumsum(1)
    Iy[1:, 1:] = dy.astype(np.float64).cumsum(0).cumsum(1)
    out = np.full((h, w, 3), np.nan, np.float32)
    b = int(smoothing)
    for r in range(b, h - b):
        for c in range(b, w - b):
            sm = min(dm[r, c], np.float32(smoothing))
            if not (np.isfinite(z[r, c]) and sm > 2.0):
                continue
            s = int(sm)
            x0, y0 = c - s // 2, r - s // 2
            gx = Ix[y0 + s, x0 + s] + Ix[y0, x0] - Ix[y0, x0 + s] - Ix[y0 + s, x0]
            gy = Iy[y0 + s, x0 + s] + Iy[y0, x0] - Iy[y0, x0 + s] - Iy[y0 + s, x0]
            n = np.cross(gy, gx)
            l2 = float(n @ n)
            if l2 == 0.0:
                continue
            n = (n / np.sqrt(l2)).astype(np.float32)
            if float(-(P[r, c] @ n)) < 0:
                n = -n
            out[r, c] = n
    return out, dm


def test_restatement_agrees_with_an_independent_formulation(drfe, orc):
    """the C++ restatement against brute_force_normals on a real-looking cloud (depth jumps, culled far points, a hole): identical
    distance map (float for float) and NaN pattern, normals equal to 1e-5 (the window sums are associated differently)"""
    from test_peac import clean_depth
    _, depth, K = drfe.synth_frame(640, 480, 2, 20260100)
    cloud = orc.third_cloud(clean_depth(depth, ((200, 260, 100, 180),)), *K, 4.0)
    nr, dm = orc.integral_normals(cloud, with_distance_map=True)
    bn, bdm = brute_force_normals(cloud)
    assert np.array_equal(dm, bdm)
    # a window whose gradients cancel to exactly 0 in one association may leave 1e-17 in the other: compare where both are finite
    both = ~np.isnan(nr[..., 0]) & ~np.isnan(bn[..., 0])
    assert (np.isnan(nr[..., 0]) != np.isnan(bn[..., 0])).mean() < 1e-3 and both.mean() > 0.2
    assert np.abs(nr[both] - bn[both]).max() < 1e-5
