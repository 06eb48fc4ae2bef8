"""PEAC-AHC, the plane extractor that is live in Frame::Frame (reference src/Frame.cc:126, :937-949; include/peac/*.hpp).
CPU: invariants of the restatement (oracle/peac_oracle.cpp).  GPU: drfe_peac_* against the restatement — seg_output, plane
order, plane_vertices_ and member points bit for bit, plane parameters bit for bit (same Jacobi solver, same operation order)."""
import numpy as np
import pytest


def clean_depth(depth, holes=()):
    """the synthetic generator drops 2 % of the pixels independently; with INIT_STRICT (no missing pixel allowed in a 10x10 window,
    the reference's default) that leaves almost no window, so the dropouts are filled from a neighbour and coherent holes are cut
    instead, like a real sensor's"""
    d = depth.copy()
    for _ in range(4):
        z = d == 0
        if not z.any():
            break
        p = np.pad(d, 1, mode="edge")
        nb = np.max(np.stack([p[1:-1, :-2], p[1:-1, 2:], p[:-2, 1:-1], p[2:, 1:-1]]), 0)
        d[z] = nb[z]
    for (y0, y1, x0, x1) in holes:
        d[y0:y1, x0:x1] = 0
    return d


def frame(drfe, scene, seed, holes=((100, 140, 300, 420),)):
    _, depth, K = drfe.synth_frame(640, 480, scene, seed)
    return np.rint(clean_depth(depth, holes) * 5000).astype(np.uint16), K


FAC = float(np.float32(1.0 / 5000.0))


def test_peac_restatement_invariants(drfe, orc):
    q, K = frame(drfe, 1, 20260012)
    cloud = orc.peac_cloud(q, FAC, *K)
    assert cloud.shape == (640 * 480, 3) and (cloud[:, 2] <= 5.0).all()
    z = q.astype(np.float64).ravel() * np.float64(np.float32(FAC))
    assert np.array_equal(cloud[:, 2], np.where(z > 5.0, 0.0, z))
    seg, planes, members, steps = orc.peac_run(cloud, 640, 480)
    n = len(planes)
    assert n >= 3 and steps > 500 and seg.max() == n
    assert np.all(np.diff(planes[:, 8]) <= 0)                                  # sorted by N, decreasing
    assert np.allclose((planes[:, :3] ** 2).sum(1), 1, atol=1e-12)
    assert np.all((planes[:, :3] * planes[:, 3:6]).sum(1) <= 0)                # normals point towards the camera
    for p in range(n):
        idx = members[p]
        assert np.all(np.diff(idx) > 0) and np.all(seg.ravel()[idx] == p + 1)  # scan order, consistent with seg_output
    assert sum(len(m) for m in members) == int((seg > 0).sum())
    assert (seg[100:140, 300:420] == 0).all()                                   # nothing grows into a hole (no depth there)
    # every member lies close to its plane (region growing accepts within 3 sigma of the block-level fit)
    for p in range(n):
        pts = cloud[members[p]]
        d = np.abs((pts - planes[p, 3:6]) @ planes[p, :3])
        assert np.median(d) < 0.02


def test_peac_fit_is_pinned_to_lapack(orc):
    """Stats::compute (AHCPlaneSeg.hpp:128-162) with the restatement's Jacobi solver in place of Eigen's SelfAdjointEigenSolver (P.1):
    smallest eigenvalue, its vector and the curvature agree with numpy.linalg.eigh (LAPACK dsyevd) on the same scatter matrix to a
    few ulps of the LARGEST eigenvalue — what any backward-stable solver, Eigen's included, delivers"""
    rng = np.random.default_rng(0)
    worst = np.zeros(3)
    for trial in range(1500):
        N = int(rng.integers(50, 60000)); ext = rng.uniform(0.02, 1.5); sig = 10 ** rng.uniform(-4, -1.5)
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        pts = np.c_[rng.uniform(-ext, ext, N), rng.uniform(-ext * rng.uniform(0.1, 1), ext, N), rng.normal(0, sig, N)] @ R.T
        pts += rng.uniform(-1, 1, 3) * np.array([1, 1, 0]) + np.array([0, 0, rng.uniform(0.5, 4)])
        s = pts.sum(0); sq = (pts * pts).sum(0)
        sxy, syz, sxz = (pts[:, 0] * pts[:, 1]).sum(), (pts[:, 1] * pts[:, 2]).sum(), (pts[:, 0] * pts[:, 2]).sum()
        ctr, nrm, mse, curv = orc.peac_fit([s[0], s[1], s[2], sq[0], sq[1], sq[2], sxy, syz, sxz], N)
        sc = 1.0 / N
        K = np.array([[sq[0] - s[0] * s[0] * sc, sxy - s[0] * s[1] * sc, sxz - s[0] * s[2] * sc],
                      [0, sq[1] - s[1] * s[1] * sc, syz - s[1] * s[2] * sc], [0, 0, sq[2] - s[2] * s[2] * sc]])
        K = K + np.triu(K, 1).T
        w, V = np.linalg.eigh(K)
        assert np.array_equal(ctr, s * sc) and nrm @ ctr <= 0                     # the normal points towards the camera
        worst = np.maximum(worst, [abs(mse / sc - w[0]) / w[2], 1 - abs(nrm @ V[:, 0]), abs(curv - w[0] / w.sum())])
    assert worst[0] < 4e-15 and worst[1] < 1e-14 and worst[2] < 4e-15, worst


@pytest.mark.parametrize("scene,seed,holes,kw", [
    (0, 20260000, ((100, 140, 300, 420),), {}),
    (2, 20260100, ((0, 60, 0, 640), (200, 260, 100, 180)), {}),
    (1, 20260777, ((200, 260, 100, 180),), dict(min_support=1500, win_w=16, win_h=12)),
])
def test_two_independent_restatements_agree(drfe, orc, scene, seed, holes, kw):
    """oracle/peac_py.py (plain Python, heapq / set / list, written from the reference headers) against oracle/peac_oracle.cpp:
    with the declared Jacobi solver every value is equal, the doubles bit for bit; with LAPACK (numpy.linalg.eigh) in the place where
    the reference has Eigen's SelfAdjointEigenSolver, the decisions are the same — same number of clustering steps, same seg_output,
    same plane order and members — and the plane parameters agree to 1e-12: the solver substitution P.1 changes nothing that is extracted"""
    from oracle import peac_py
    q, K = frame(drfe, scene, seed, holes)
    cloud = orc.peac_cloud(q, FAC, *K)
    assert np.array_equal(cloud, peac_py.cloud_from_depth(q, FAC, *K))
    ms, ww, wh = kw.get("min_support", 3000), kw.get("win_w", 10), kw.get("win_h", 10)
    oseg, oplanes, omem, osteps = orc.peac_run(cloud, 640, 480, min_support=ms, window=(ww, wh))
    assert len(oplanes) >= 3
    for solver in ("jacobi", "lapack"):
        seg, planes, mem, steps = peac_py.run(cloud, 640, 480, min_support=ms, win_w=ww, win_h=wh, solver=solver)
        assert steps == osteps and np.array_equal(seg, oseg) and len(planes) == len(oplanes), solver
        assert all(np.array_equal(a, b) for a, b in zip(mem, omem)), solver
        if solver == "jacobi":
            assert planes.tobytes() == oplanes[:, :10].tobytes()
        else:
            assert np.array_equal(planes[:, 8:], oplanes[:, 8:10]) and np.abs(planes[:, :8] - oplanes[:, :8]).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("scene,seed,holes", [
    (0, 20260000, ((100, 140, 300, 420),)),
    (1, 20260012, ((100, 140, 300, 420),)),
    (2, 20260100, ((0, 60, 0, 640), (200, 260, 100, 180))),
    (1, 20260042, ()),
])
def test_gpu_peac_parity(drfe, orc, scene, seed, holes):
    q, K = frame(drfe, scene, seed, holes)
    pe = drfe.PEAC(640, 480)
    pe.enqueue(q[None], FAC, *K)
    seg, planes, npl = pe.download()
    idx, pts, offs = pe.plane_vertices()
    cloud = orc.peac_cloud(q, FAC, *K)
    oseg, oplanes, omem, osteps = orc.peac_run(cloud, 640, 480, params=pe.params_array())
    cnt = pe.counters(0)
    assert cnt[0] == osteps, (cnt, osteps)
    assert npl[0] == len(oplanes)
    n = int(npl[0])
    assert np.array_equal(planes[0, :n]["N"], oplanes[:, 8].astype(np.int32)) and np.array_equal(planes[0, :n]["rid"], oplanes[:, 9].astype(np.int32))
    assert np.array_equal(planes[0, :n]["normal"], oplanes[:, :3]) and np.array_equal(planes[0, :n]["center"], oplanes[:, 3:6])
    assert np.array_equal(planes[0, :n]["mse"], oplanes[:, 6]) and np.array_equal(planes[0, :n]["curvature"], oplanes[:, 7])
    assert np.array_equal(seg[0], oseg)
    for p in range(n):
        got = idx[0, offs[0, p]:offs[0, p + 1]]
        assert np.array_equal(got, omem[p]), p
        assert np.array_equal(pts[0, offs[0, p]:offs[0, p + 1]], cloud[omem[p]].astype(np.float32)), p


@pytest.mark.gpu
def test_gpu_peac_batch(drfe, orc):
    B = 10
    fr = [frame(drfe, i % 3, 20260600 + 3 * i, ((40 * (i % 5), 40 * (i % 5) + 50, 50 * i, 50 * i + 90),)) for i in range(B)]
    q = np.stack([f[0] for f in fr]); K = fr[0][1]
    pe = drfe.PEAC(640, 480, max_batch=B)
    pe.enqueue(q, FAC, *K)
    seg, planes, npl = pe.download()
    for f in range(B):
        oseg, oplanes, omem, _ = orc.peac_run(orc.peac_cloud(q[f], FAC, *K), 640, 480)
        assert npl[f] == len(oplanes) and np.array_equal(seg[f], oseg), f
        assert np.array_equal(planes[f, :npl[f]]["normal"], oplanes[:, :3]), f
    pe1 = drfe.PEAC(640, 480)                                   # a frame alone = the same frame in a batch
    pe1.enqueue(q[7:8], FAC, *K)
    s1, p1, n1 = pe1.download()
    assert n1[0] == npl[7] and np.array_equal(s1[0], seg[7])
    with pytest.raises(drfe.DrfeError):
        pe1.download(plane_cap=1)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [
    dict(window_width=16, window_height=12, min_support=1500),                     # non-square windows (40 x 40 grid), smaller planes kept
    dict(window_width=20, window_height=20, min_support=5000),                     # the AHC paper's coarse setting
    dict(similarityTh_merge=float(np.cos(np.pi / 12)), similarityTh_refine=float(np.cos(np.pi / 18)), stdTol_merge=0.004, depthAlpha=0.02),
    dict(max_depth=2.5),                                                            # a tighter readDepthImage cull: more of the frame is empty
])
def test_gpu_peac_other_parameters(drfe, orc, kw):
    """ahc::ParamSet / PlaneFitter members other than DR-SLAM's defaults: same parity"""
    q, K = frame(drfe, 1, 20260777, ((200, 260, 100, 180),))
    pe = drfe.PEAC(640, 480, **kw)
    pe.enqueue(q[None], FAC, *K)
    seg, planes, npl = pe.download()
    idx, pts, offs = pe.plane_vertices()
    cloud = orc.peac_cloud(q, FAC, *K)
    if "max_depth" in kw:                                                          # the oracle's cloud culls at 5.0 (PlaneExtractor.cpp:44): cull again
        cloud = cloud.copy(); cloud[cloud[:, 2] > kw["max_depth"]] = 0.0
    oseg, oplanes, omem, osteps = orc.peac_run(cloud, 640, 480, params=pe.params_array(), min_support=kw.get("min_support", 3000),
                                               window=(kw.get("window_width", 10), kw.get("window_height", 10)))
    assert pe.counters(0)[0] == osteps and npl[0] == len(oplanes) and len(oplanes) >= 1
    n = int(npl[0])
    assert np.array_equal(seg[0], oseg)
    assert np.array_equal(planes[0, :n]["normal"], oplanes[:, :3]) and np.array_equal(planes[0, :n]["mse"], oplanes[:, 6])
    for p in range(n):
        assert np.array_equal(idx[0, offs[0, p]:offs[0, p + 1]], omem[p]), p


@pytest.mark.gpu
def test_gpu_peac_small_image_and_empty_frames(drfe, orc):
    """320 x 240 with its own intrinsics; a frame without depth and a frame of pure noise give no plane and an all-zero seg_output"""
    _, depth, K = drfe.synth_frame(320, 240, 2, 20260801)
    q = np.rint(clean_depth(depth) * 5000).astype(np.uint16)
    rng = np.random.default_rng(5)
    batch = np.stack([q, np.zeros_like(q), rng.integers(500, 20000, q.shape).astype(np.uint16)])
    pe = drfe.PEAC(320, 240, max_batch=3, min_support=800)
    pe.enqueue(batch, FAC, *K)
    seg, planes, npl = pe.download()
    idx, pts, offs = pe.plane_vertices()
    for f in range(3):
        oseg, oplanes, omem, _ = orc.peac_run(orc.peac_cloud(batch[f], FAC, *K), 320, 240, params=pe.params_array(), min_support=800)
        assert npl[f] == len(oplanes) and np.array_equal(seg[f], oseg), f
        for p in range(len(oplanes)):
            assert np.array_equal(idx[f, offs[f, p]:offs[f, p + 1]], omem[p])
    assert npl[0] >= 1 and npl[1] == 0 and not seg[1].any() and offs[1].max() == 0
    with pytest.raises(drfe.DrfeError):                                            # 80 x 60 windows: more than the clustering kernel's 4096
        drfe.PEAC(640, 480, window_width=8, window_height=8)


def _golden_frames():
    from conftest import load_golden
    G = load_golden("peac_640x480.npz")
    for i in range(2):
        offs = G["member_offsets%d" % i]
        yield (G["depth%d" % i], [float(v) for v in G["K%d" % i]], float(G["depth_factor"]), G["seg%d" % i], G["planes%d" % i],
               [G["member_idx%d" % i][offs[p]:offs[p + 1]] for p in range(len(offs) - 1)], int(G["steps%d" % i]))


def test_oracle_reproduces_peac_golden(orc):
    """tests/golden/peac_640x480.npz (made by tests/golden/make_golden_peac.py): the restatement must not drift"""
    for q, K, fac, seg, planes, members, steps in _golden_frames():
        oseg, oplanes, omem, osteps = orc.peac_run(orc.peac_cloud(q, fac, *K), 640, 480)
        assert osteps == steps and np.array_equal(oseg, seg) and oplanes.tobytes() == planes.tobytes()
        assert len(omem) == len(members) and all(np.array_equal(a, b) for a, b in zip(omem, members))


@pytest.mark.gpu
def test_gpu_reproduces_peac_golden(drfe):
    for q, K, fac, seg, planes, members, steps in _golden_frames():
        pe = drfe.PEAC(640, 480)
        pe.enqueue(q[None], fac, *K)
        gseg, gplanes, npl = pe.download()
        idx, pts, offs = pe.plane_vertices()
        n = int(npl[0])
        assert n == len(planes) and pe.counters(0)[0] == steps and np.array_equal(gseg[0], seg)
        for name, cols in (("normal", slice(0, 3)), ("center", slice(3, 6))):
            assert np.array_equal(gplanes[0, :n][name], planes[:, cols]), name
        assert np.array_equal(gplanes[0, :n]["mse"], planes[:, 6]) and np.array_equal(gplanes[0, :n]["curvature"], planes[:, 7])
        assert np.array_equal(gplanes[0, :n]["N"], planes[:, 8].astype(np.int32)) and np.array_equal(gplanes[0, :n]["rid"], planes[:, 9].astype(np.int32))
        for p in range(n):
            assert np.array_equal(idx[0, offs[0, p]:offs[0, p + 1]], members[p]), p


@pytest.mark.gpu
def test_gpu_peac_argument_and_state_errors(drfe):
    """the C ABI's error behaviour: every misuse comes back as a status with a text, nothing is launched"""
    import ctypes as C
    L = drfe.lib()
    pe = drfe.PEAC(640, 480, max_batch=2)
    seg = np.zeros((1, 480, 640), np.uint8); planes = np.zeros((1, 255), drfe.PEAC_PLANE_DTYPE); npl = np.zeros(1, np.int32)
    assert L.drfe_peac_download(pe.h, seg.ctypes.data, planes.ctypes.data, 255, npl.ctypes.data) == drfe.ERR_STATE      # nothing enqueued
    q = np.zeros((3, 480, 640), np.uint16)
    for nframes, rs, kind in ((3, 640, drfe.MEM_HOST), (0, 640, drfe.MEM_HOST), (1, 600, drfe.MEM_HOST), (1, 640, 7)):
        rc = L.drfe_peac_enqueue_depth_u16(pe.h, nframes, q.ctypes.data, rs, rs * 480, kind, C.c_float(2e-4), C.c_float(525), C.c_float(525),
                                           C.c_float(320), C.c_float(240))
        assert rc == drfe.ERR_ARG and len(L.drfe_last_error()) > 10, (nframes, rs, kind)
    assert L.drfe_peac_enqueue_depth_u16(None, 1, q.ctypes.data, 640, 640 * 480, drfe.MEM_HOST, C.c_float(2e-4), C.c_float(525), C.c_float(525),
                                         C.c_float(320), C.c_float(240)) == drfe.ERR_ARG
    prm = drfe.PeacParams()
    L.drfe_peac_default_params(C.byref(prm))
    h = C.c_void_p()
    assert L.drfe_peac_create(640, 480, C.byref(prm), 0, 0, C.byref(h)) == drfe.ERR_ARG                                   # max_batch 0
    prm.window_width = 0
    assert L.drfe_peac_create(640, 480, C.byref(prm), 1, 0, C.byref(h)) == drfe.ERR_ARG
    pe.enqueue(q[:2], 2e-4, 525.0, 525.0, 320.0, 240.0)                                                                    # still usable
    assert pe.download()[2].tolist() == [0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("max_dist", [3.0, float(np.finfo(np.float32).max)])
def test_gpu_peac_plane_points_voxel(drfe, orc, max_dist):
    """Frame::ComputePlanes' per-plane clouds (Frame.cc:954-990): vertices with (float) z <= mMax_point_dist, then the 5 cm pcl::VoxelGrid —
    drfe_peac_plane_points_voxel against the member lists of the PEAC restatement pushed through the voxel-grid restatement"""
    fr = [frame(drfe, s, seed, holes) for s, seed, holes in ((1, 20260012, ((100, 140, 300, 420),)), (2, 20260100, ((200, 260, 100, 180),)))]
    q = np.stack([f[0] for f in fr]); K = fr[0][1]
    pe = drfe.PEAC(640, 480, max_batch=2)
    pe.enqueue(q, FAC, *K)
    seg, planes, npl = pe.download()
    pts, offs = pe.plane_points_voxel(max_dist, 0.05)
    for f in range(2):
        cloud = orc.peac_cloud(q[f], FAC, *K)
        _, oplanes, omem, _ = orc.peac_run(cloud, 640, 480)
        assert npl[f] == len(oplanes) and len(oplanes) >= 3
        culled = 0
        for p in range(len(oplanes)):
            xyz = cloud[omem[p]].astype(np.float32)
            keep = ~(xyz[:, 2] > np.float32(max_dist))
            culled += int((~keep).sum())
            want, unfiltered = orc.voxel_grid(xyz[keep], 0.05)
            got = pts[f, offs[f, p]:offs[f, p + 1]]
            assert not unfiltered and got.shape == want.shape and got.tobytes() == want.tobytes(), (f, p)
        assert (culled > 0) == (max_dist < 100)                   # the 3 m cull removes something on these scenes
