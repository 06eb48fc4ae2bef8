"""ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (reference
src/ORBmatcher.cc:1396-1535), the matcher TrackWithMotionModel runs every frame.  CPU: the oracle's literal sequential
loop against properties a brute-force check can state (projection model = cv2.gemm, every match is the first minimum over
the brute-force window, no observed keypoint is taken twice).  GPU: drfe_orb_search_last_frame (parallel sweeps to the
fixed point, C ABI) against the oracle — every output bit-exact, conflicts and cascades included."""
import numpy as np
import pytest

K = (525.0, 525.0, 319.5, 239.5)
DIST = [0.1, -0.05, 0.001, 0.0005, 0.0]


def small_pose(rng, rot_deg=1.0, trans=0.03):
    """Tcw (3x4 float32) of a small motion"""
    w = rng.normal(0, np.deg2rad(rot_deg), 3)
    th = np.linalg.norm(w)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    t = rng.normal(0, trans, 3)
    return np.hstack([R, t[:, None]]).astype(np.float32)


def make_last_frame(orc, p, keys_un, kp_depth, desc, Tcw, npts, seed, cascade=True):
    """A last frame whose map points project near the current frame's keypoints: world position = the keypoint
    back-projected at its depth (+ noise) through Tcw^-1, octave near the keypoint's, descriptor = the keypoint's with a
    few bits flipped.  ~12 % of the points aim at a keypoint another point already aims at (conflicts), a block of points
    shares one target neighbourhood and one descriptor (cascades: each one is pushed to its next-best keypoint)."""
    rng = np.random.default_rng(seed)
    n = len(keys_un)
    pts = np.zeros(npts, orc.LAST_POINT_DTYPE)
    pd = np.zeros((npts, 32), np.uint8)
    src = rng.integers(0, n, npts)
    dup = rng.random(npts) < 0.12
    src[dup] = src[rng.integers(0, npts, dup.sum())]
    if cascade:
        # 24 points aiming at the same place with the same descriptor
        src[40:64] = src[40]
    z = np.where(kp_depth[src] > 0, kp_depth[src], 2.0).astype(np.float64)
    u = keys_un["x"][src] + rng.normal(0, 2.0, npts)
    v = keys_un["y"][src] + rng.normal(0, 2.0, npts)
    Xc = np.stack([(u - p.cx) * z / p.fx, (v - p.cy) * z / p.fy, z], 1)
    R, t = Tcw[:, :3].astype(np.float64), Tcw[:, 3].astype(np.float64)
    Xw = (Xc - t) @ R                                  # R^T (Xc - t)
    pts["X"], pts["Y"], pts["Z"] = Xw[:, 0], Xw[:, 1], Xw[:, 2]
    pts["octave"] = np.clip(keys_un["octave"][src] + rng.integers(-1, 2, npts), 0, 7)
    ang = keys_un["angle"][src] + rng.normal(0, 4, npts) + np.where(rng.random(npts) < 0.15, rng.uniform(0, 360, npts), 0)
    pts["angle"] = np.mod(ang, 360).astype(np.float32)
    pts["flags"] = orc.LP_VALID * (rng.random(npts) < 0.93) + orc.LP_OBSERVED * (rng.random(npts) < 0.8)
    pd[:] = desc[src]
    flips = rng.integers(0, 256, (npts, 40))
    nfl = rng.integers(0, 40, npts)
    for i in range(npts):
        for b in flips[i, :nfl[i]]:
            pd[i, b >> 3] ^= 1 << (b & 7)
    if cascade:
        pd[40:64] = pd[40]
        pts["flags"][40:64] = orc.LP_VALID | orc.LP_OBSERVED
        pts["flags"][44] = orc.LP_VALID                # one temporal point (no observations) inside the cascade
    # behind the camera / outside the image / exactly on the camera plane
    pts["Z"][5] = -abs(pts["Z"][5]) - 5
    pts["X"][6] += 50
    return pts, pd


def current_frame(drfe, orc, seed, scene=1, size=(640, 480), nfeatures=1000):
    gray, depth, _ = drfe.synth_frame(size[0], size[1], scene, seed)
    o = orc.OrbOracle(nfeatures)
    keys, desc = o.extract(gray)
    p = orc.frame_params(*K, DIST, 40.0, size[0], size[1])
    ku, ur, kd, gc, gi = orc.frame_post(p, keys, depth)
    return gray, depth, p, ku, ur, kd, gc, gi, desc, np.array(o.scale_factors(), np.float32)


def test_projection_model_is_cv2_gemm():
    """x3Dc = Rcw*x3Dw + tcw is a cv::MatExpr that evaluates as cv::gemm(Rcw, x3Dw, 1, tcw, 1): products and sums in float,
    left to right, then the addition of tcw — the model the oracle and the kernel use"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    f32 = np.float32
    for _ in range(3000):
        R = rng.normal(0, 1, (3, 3)).astype(f32)
        X = rng.normal(0, 3, (3, 1)).astype(f32)
        c = rng.normal(0, 2, (3, 1)).astype(f32)
        got = cv2.gemm(R, X, 1, c, 1)
        for r in range(3):
            t0 = f32(f32(f32(R[r, 0] * X[0, 0]) + f32(R[r, 1] * X[1, 0])) + f32(R[r, 2] * X[2, 0]))
            assert got[r, 0] == f32(np.float64(t0) + np.float64(c[r, 0]))


@pytest.mark.parametrize("mode,check", [(0, 1), (1, 1), (2, 0)])
def test_oracle_last_frame_properties(drfe, orc, mode, check):
    gray, depth, p, ku, ur, kd, gc, gi, desc, sf = current_frame(drfe, orc, 20260430)
    n = len(ku)
    rng = np.random.default_rng(1)
    Tcw = small_pose(rng)
    pts, pd = make_last_frame(orc, p, ku, kd, desc, Tcw, 700, 3)
    occ = (rng.random(n) < 0.05).astype(np.uint8)
    th = 15.0
    mk, md, holder, nm = orc.search_last_frame(p, sf, ku, ur, gc, gi, desc, Tcw.ravel(), th, mode, check, pts, pd, occ)
    matched = np.nonzero(mk >= 0)[0]
    assert len(matched) > 300
    # brute force: projection in float64, windows without the grid; candidate sets can differ from the float32 walk only for
    # keypoints within rounding of the window edge, so compare where the margin is clear
    placed = np.zeros(n, bool)
    placed[gi] = True
    ham = np.unpackbits(pd[:, None, :] ^ desc[None, :, :], axis=2).sum(2)
    taken = occ.astype(bool).copy()
    checked = 0
    for i in range(len(pts)):
        lp = pts[i]
        if not lp["flags"] & orc.LP_VALID:
            assert mk[i] == -1
            continue
        Xc = Tcw[:, :3].astype(np.float64) @ np.array([lp["X"], lp["Y"], lp["Z"]], np.float64) + Tcw[:, 3]
        if Xc[2] <= 0:
            assert mk[i] == -1
            continue
        u, v = p.fx * Xc[0] / Xc[2] + p.cx, p.fy * Xc[1] / Xc[2] + p.cy
        r = th * sf[lp["octave"]]
        o = lp["octave"]
        lo, hi = (o, 99) if mode == 1 else (0, o) if mode == 2 else (o - 1, o + 1)
        dx, dy = np.abs(ku["x"] - u), np.abs(ku["y"] - v)
        ok = placed & (dx < r) & (dy < r) & (ku["octave"] >= lo) & (ku["octave"] <= hi) & ~taken
        urp = u - p.bf / Xc[2]
        er = np.abs(urp - ur)
        ok &= ~((ur > 0) & (er > r))
        edge = (placed & ((np.abs(dx - r) < 1e-2) | (np.abs(dy - r) < 1e-2) | ((ur > 0) & (np.abs(er - r) < 1e-2)))).any()
        inside = p.min_x + 1e-2 < u < p.max_x - 1e-2 and p.min_y + 1e-2 < v < p.max_y - 1e-2
        if not inside and not (p.min_x - 1e-2 < u < p.max_x + 1e-2 and p.min_y - 1e-2 < v < p.max_y + 1e-2):
            assert mk[i] == -1
        if inside and not edge:
            d = np.where(ok, ham[i], 999)
            checked += 1
            if ok.any() and d.min() <= orc.TH_HIGH:
                assert mk[i] >= 0 and md[i] == d.min() and d[mk[i]] == d.min(), i
            else:
                assert mk[i] == -1, i
        if mk[i] >= 0 and lp["flags"] & orc.LP_OBSERVED:
            taken[mk[i]] = True
    assert checked > 500
    # no keypoint is taken by two observed points; the cascade block spread over distinct keypoints
    obs = matched[(pts["flags"][matched] & orc.LP_OBSERVED) != 0]
    assert len(set(mk[obs])) == len(obs)
    block = [k for k in mk[40:64] if k >= 0]
    assert len(block) >= 2 and len(set(block)) >= len(block) - 1      # only the temporal point's keypoint may be re-taken
    # holder / nmatches bookkeeping
    removed = (holder == -2).sum()
    if not check:
        assert removed == 0 and nm == len(matched)
    else:
        assert nm <= len(matched) and removed > 0
    for idx in np.nonzero(holder >= 0)[0]:
        assert mk[holder[idx]] == idx


def sweeps_model(orc, p, sf, ku, ur, gc, gi, desc, Tcw, th, mode, pts, pd, occ):
    """numpy model of k_search_last_frame's fixed-point iteration (what the kernel does, minus the rotation check):
    candidate lists in GetFeaturesInArea order, then sweeps with owner[idx] = lowest observed point that chose idx"""
    f32, f64 = np.float32, np.float64
    T = np.asarray(Tcw, f32).reshape(3, 4)
    off = np.concatenate([[0], np.cumsum(np.asarray(gc).ravel())]).astype(np.int64)
    d32 = np.ascontiguousarray(desc).view(np.uint32).reshape(len(desc), 8)
    cands = []
    for i, lp in enumerate(pts):
        c = []
        if lp["flags"] & orc.LP_VALID:
            X = (f32(lp["X"]), f32(lp["Y"]), f32(lp["Z"]))
            c3 = [f32(f64(f32(f32(f32(T[r, 0] * X[0]) + f32(T[r, 1] * X[1])) + f32(T[r, 2] * X[2]))) + f64(T[r, 3])) for r in range(3)]
            invzc = f32(f64(1.0) / f64(c3[2]))
            u = f32(f32(f32(f32(p.fx) * c3[0]) * invzc) + f32(p.cx))
            v = f32(f32(f32(f32(p.fy) * c3[1]) * invzc) + f32(p.cy))
            if not (invzc < 0) and p.min_x <= u <= p.max_x and p.min_y <= v <= p.max_y:
                o = int(lp["octave"])
                radius = f32(f32(th) * f32(sf[o]))
                lo, hi = (o, -1) if mode == 1 else (0, o) if mode == 2 else (o - 1, o + 1)
                urp = f32(u - f32(f32(p.bf) * invzc))
                q = np.ascontiguousarray(pd[i]).view(np.uint32)
                for idx in orc.features_in_area(p, ku, off, gi, u, v, radius, lo, hi):
                    if ur[idx] > 0 and abs(f32(urp - ur[idx])) > radius:
                        continue
                    c.append((idx, orc.descriptor_distance(q, d32[idx])))
        cands.append(c)
    choice = np.full(len(pts), -1, np.int64)
    sweeps = 0
    while True:
        owner = np.where(occ != 0, -1, 2 ** 31 - 1).astype(np.int64)
        for i in range(len(pts)):
            if choice[i] >= 0 and pts["flags"][i] & orc.LP_OBSERVED:
                owner[choice[i]] = min(owner[choice[i]], i)
        new = np.full(len(pts), -1, np.int64)
        for i, c in enumerate(cands):
            best, bi = 256, -1
            for idx, d in c:
                if owner[idx] < i:
                    continue
                if d < best:
                    best, bi = d, idx
            new[i] = bi if best <= orc.TH_HIGH else -1
        sweeps += 1
        if np.array_equal(new, choice):
            return choice, sweeps
        choice = new


def test_parallel_sweeps_reach_the_sequential_result(drfe, orc):
    """the design claim of k_search_last_frame: the fixed point of the parallel sweeps is the reference's in-order result"""
    gray, depth, p, ku, ur, kd, gc, gi, desc, sf = current_frame(drfe, orc, 20260431, scene=2)
    rng = np.random.default_rng(4)
    for mode in (0, 2):
        Tcw = small_pose(rng)
        pts, pd = make_last_frame(orc, p, ku, kd, desc, Tcw, 600, 8 + mode)
        occ = (rng.random(len(ku)) < 0.05).astype(np.uint8)
        mk, md, holder, nm = orc.search_last_frame(p, sf, ku, ur, gc, gi, desc, Tcw.ravel(), 15.0, mode, 0, pts, pd, occ)
        choice, sweeps = sweeps_model(orc, p, sf, ku, ur, gc, gi, desc, Tcw.ravel(), 15.0, mode, pts, pd, occ)
        assert np.array_equal(choice, mk)
        assert 3 <= sweeps < 40


def gpu_case(drfe, orc, seeds, scenes, modes, checks, ths, npts, with_occ, size=(640, 480), nfeatures=1000):
    B = len(seeds)
    frames = [current_frame(drfe, orc, s, scene=sc, size=size, nfeatures=nfeatures) for s, sc in zip(seeds, scenes)]
    ex = drfe.ORBextractor(nfeatures, 1.2, 8, 20, 7, size[0], size[1], max_batch=B)
    ex.enqueue(np.stack([f[0] for f in frames]))
    kps, desc, cnt = ex.download()
    p = ex.frame_params(*K, DIST, 40.0)
    ku, ur, kd, gc, gi = ex.frame_post(p, np.stack([f[1] for f in frames]))
    pcap = max(npts)
    tp = np.zeros(B, drfe.TRACK_PARAMS_DTYPE)
    P = np.zeros((B, pcap), drfe.LAST_POINT_DTYPE)
    PD = np.zeros((B, pcap, 32), np.uint8)
    rng = np.random.default_rng(77)
    occ = (rng.random((B, ex.cap)) < 0.05).astype(np.uint8) if with_occ else None
    for f in range(B):
        n = int(cnt[f])
        assert np.array_equal(desc[f, :n], frames[f][8])
        Tcw = small_pose(rng)
        tp[f]["Tcw"], tp[f]["th"], tp[f]["mode"], tp[f]["check_orientation"] = Tcw.ravel(), ths[f], modes[f], checks[f]
        if npts[f]:
            pts, pd = make_last_frame(orc, p, ku[f, :n], kd[f, :n], desc[f, :n], Tcw, npts[f], 100 + f, cascade=npts[f] > 64)
            P[f, :npts[f]], PD[f, :npts[f]] = pts, pd
    mk, md, kp, nm, sw = ex.search_last_frame(tp, P, PD, np.array(npts, np.int32), occ)
    for f in range(B):
        n, m = int(cnt[f]), npts[f]
        placed = int(gc[f].sum())
        wmk, wmd, wh, wnm = orc.search_last_frame(frames[f][2], frames[f][9], ku[f, :n], ur[f, :n], gc[f], gi[f, :placed], desc[f, :n],
                                                  tp[f]["Tcw"], ths[f], modes[f], checks[f], P[f, :m], PD[f, :m],
                                                  None if occ is None else occ[f, :n])
        assert np.array_equal(mk[f, :m], wmk), f
        assert np.array_equal(md[f, :m], wmd), f
        assert np.array_equal(kp[f, :n], wh), f
        assert nm[f] == wnm, f
        assert (kp[f, n:] == -1).all() and (mk[f, m:] == -1).all() and (md[f, m:] == 256).all()
    return mk, sw, nm


@pytest.mark.gpu
def test_gpu_search_last_frame(drfe, orc):
    mk, sw, nm = gpu_case(drfe, orc, [20260430, 20260435, 20260440, 20260445], [1, 0, 2, 1], [0, 1, 2, 0], [1, 1, 0, 1],
                          [15.0, 7.0, 30.0, 15.0], [900, 700, 500, 0], True)
    assert (mk[0] >= 0).sum() > 400 and nm[3] == 0
    assert sw[0] >= 3                    # the cascade block needs more than two sweeps
    assert sw[3] == 1


@pytest.mark.gpu
def test_gpu_search_last_frame_no_occupied(drfe, orc):
    """what TrackWithMotionModel does: mvpMapPoints filled with NULL before the call"""
    gpu_case(drfe, orc, [20260450, 20260455], [1, 2], [0, 0], [1, 1], [15.0, 15.0], [1000, 30], False)


@pytest.mark.gpu
def test_gpu_search_last_frame_1280x720_2000kp(drfe, orc):
    """BASELINE configs[4] shape: 2000 keypoints per frame, 2400 last-frame points (capacities beyond one CTA pass)"""
    mk, sw, nm = gpu_case(drfe, orc, [20260460, 20260461], [2, 1], [0, 2], [1, 1], [15.0, 7.0], [2400, 1500], True, size=(1280, 720), nfeatures=2000)
    assert (mk[0] >= 0).sum() > 1000
